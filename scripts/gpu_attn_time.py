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
us = t(lambda: ops.attention(q, k, vt, out, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S, scale=scale, qext=qext, kext=kext, row_bias=rb, ext_cols=64))
fl = 4 * B * H * S * S * hd
print(f"global attention B={B}: {us:8.1f} us  {fl / us / 1e6:6.0f} TF/s  (LLMSEG_ATTN_V2={os.environ.get('LLMSEG_ATTN_V2', '0')})", flush=True)
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
