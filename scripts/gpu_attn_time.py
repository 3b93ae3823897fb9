"""Time the SAM global / window attention kernels at the batch-8 shapes (CUDA events, best of 3 x 10)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"; torch.manual_seed(0)
H, hd = 16, 80; scale = hd ** -0.5
def t(fn, n=10):
    best = 1e9
    for _ in range(3):
        fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
B = 8
S = 4096
q = torch.randn(B * H, S, hd, device=dev).bfloat16(); k = torch.randn_like(q); vt = torch.randn(B * H, hd, S, device=dev).bfloat16()
rel = ops.make_rel_hw((torch.randn(127, hd, device=dev) * 0.1).bfloat16(), (torch.randn(127, hd, device=dev) * 0.1).bfloat16())
qext = torch.zeros(B * H, S, 64, device=dev, dtype=torch.bfloat16); rb = torch.zeros_like(qext)
ops.relpos_prep(q, rel, bh=B * H, seq=S, seq_pad=S, head_dim=hd, grid=64, inv_scale=1 / scale, qext=qext, row_bias=rb)
out = torch.empty(B * S, H * hd, device=dev, dtype=torch.bfloat16); kext = ops.make_kext(64, dev)
fl = 4 * B * H * S * S * hd
outs = {}
def variant(v):   # v1: single 128-key score buffer, 8 softmax warps; v3: two 64-key buffers, 4 softmax warps
    os.environ["LLMSEG_ATTN_V1"] = "1" if v == "v1" else "0"
for v in ("v1", "v3", "v1", "v3"):      # interleaved
    variant(v)
    o = torch.empty_like(out)
    us = t(lambda: ops.attention(q, k, vt, o, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S, scale=scale, qext=qext, kext=kext, row_bias=rb, ext_cols=64))
    outs[v] = o
    print(f"global attention B={B} {v}: {us:8.1f} us  {fl / us / 1e6:6.0f} TF/s", flush=True)
print(f"global attention max|d| v3-v1 {(outs['v3'].float() - outs['v1'].float()).abs().max().item():.5f}", flush=True)
os.environ.pop("LLMSEG_ATTN_V1")
if len(sys.argv) > 1 and sys.argv[1] == "global":
    sys.exit(0)
# LLaMA causal (T=319, hd 128), CLIP (257, hd 64), DINOv2 (4097, hd 64) through both kernels
for name, nb, Hh, hdd, Sx, causal in (("llama causal T=319", B, 32, 128, 319, True), ("clip 257", B, 16, 64, 257, False),
                                      ("dinov2 4097", B, 16, 64, 4097, False), ("llama causal T=767", 2, 32, 128, 767, True)):
    Sp = (Sx + 7) // 8 * 8
    qq = torch.randn(nb * Hh, Sp, hdd, device=dev).bfloat16(); kk = torch.randn_like(qq); vv = torch.randn(nb * Hh, hdd, Sp, device=dev).bfloat16()
    res = {}
    for v in ("v1", "v3"):
        variant(v)
        oo = torch.zeros(nb * Sx, Hh * hdd, device=dev, dtype=torch.bfloat16)
        us = t(lambda: ops.attention(qq, kk, vv, oo, batch=nb, heads=Hh, head_dim=hdd, seq=Sx, seq_pad=Sp, scale=hdd ** -0.5, causal=causal))
        res[v] = (us, oo)
    print(f"{name:22s} v1 {res['v1'][0]:7.1f} us   v3 {res['v3'][0]:7.1f} us   max|d| {(res['v3'][1].float() - res['v1'][1].float()).abs().max().item():.5f}", flush=True)
os.environ.pop("LLMSEG_ATTN_V1")
S, Sp, nb = 196, 200, 25 * B
q = torch.randn(nb * H, Sp, hd, device=dev).bfloat16(); k = torch.randn_like(q); vt = torch.randn(nb * H, hd, Sp, device=dev).bfloat16()
rel = ops.make_rel_hw((torch.randn(27, hd, device=dev) * 0.1).bfloat16(), (torch.randn(27, hd, device=dev) * 0.1).bfloat16())
qext = torch.zeros(nb * H, Sp, 32, device=dev, dtype=torch.bfloat16)
ops.relpos_prep(q, rel, bh=nb * H, seq=S, seq_pad=Sp, head_dim=hd, grid=14, inv_scale=1 / scale, qext=qext)
out = torch.empty(nb * S, H * hd, device=dev, dtype=torch.bfloat16); kext = ops.make_kext(14, dev)
us = t(lambda: ops.attention(q, k, vt, out, batch=nb, heads=H, head_dim=hd, seq=S, seq_pad=Sp, scale=scale, qext=qext, kext=kext, ext_cols=32))
print(f"window attention nb={nb}: {us:8.1f} us", flush=True)
us = t(lambda: ops.relpos_prep(q, rel, bh=nb * H, seq=S, seq_pad=Sp, head_dim=hd, grid=14, inv_scale=1 / scale, qext=qext))
print(f"window relpos_prep: {us:8.1f} us", flush=True)
