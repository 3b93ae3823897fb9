"""Developer GPU check for gemm_qkv + relpos_prep + attention (run under gpurun)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops, _lib
torch.manual_seed(0)
dev = "cuda"

def rope_tables(T, hd):
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2, dtype=torch.float32, device=dev) / hd))
    fr = torch.outer(torch.arange(T, dtype=torch.float32, device=dev), inv)
    return fr.cos().bfloat16().contiguous(), fr.sin().bfloat16().contiguous()

def run(name, B, H, hd, S, causal=False, kv_len=None, grid=0, rope=False, time_it=False):
    D = H * hd
    S_pad = (S + 7) // 8 * 8
    x = torch.randn(B * S, D, device=dev).bfloat16()
    w = (torch.randn(3 * D, D, device=dev) / D ** 0.5).bfloat16()
    bias = None if rope else (torch.randn(3 * D, device=dev) * 0.1).bfloat16()
    q = torch.zeros(B * H, S_pad, hd, device=dev, dtype=torch.bfloat16)
    k = torch.zeros_like(q)
    vt = torch.zeros(B * H, hd, S_pad, device=dev, dtype=torch.bfloat16)
    cos = sin = None
    if rope: cos, sin = rope_tables(S, hd)
    ops.gemm_qkv(x, w, bias, q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=S_pad, rope_cos=cos, rope_sin=sin)
    torch.cuda.synchronize()
    # reference qkv
    ref = x.float() @ w.float().T
    if bias is not None: ref = ref + bias.float()
    ref = ref.bfloat16().float().reshape(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)  # 3,B,H,S,hd
    rq, rk, rv = ref[0], ref[1], ref[2]
    if rope:
        c = torch.cat([cos, cos], -1).float()[None, None]; s_ = torch.cat([sin, sin], -1).float()[None, None]
        def rot(t):
            h2 = hd // 2
            return torch.cat([-t[..., h2:], t[..., :h2]], -1)
        def ap(t):
            return ((t * c).bfloat16().float() + (rot(t) * s_).bfloat16().float()).bfloat16().float()
        rq, rk = ap(rq), ap(rk)
    eq = (q[:, :S].float().reshape(B, H, S, hd) - rq).abs().max().item()
    ek = (k[:, :S].float().reshape(B, H, S, hd) - rk).abs().max().item()
    ev = (vt[:, :, :S].float().reshape(B, H, hd, S).transpose(-1, -2) - rv).abs().max().item()
    print(f"[{name}] qkv split err q={eq:.4f} k={ek:.4f} v={ev:.4f}", flush=True)

    scale = hd ** -0.5
    qf = q[:, :S].float(); kf = k[:, :S].float(); vf = vt[:, :, :S].float().transpose(-1, -2)
    scores = (qf @ kf.transpose(-1, -2)) * scale
    qext = kext = rb = None; ext = 0
    if grid:
        T = 2 * grid - 1
        rel_h = (torch.randn(T, hd, device=dev) * 0.3).bfloat16(); rel_w = (torch.randn(T, hd, device=dev) * 0.3).bfloat16()
        idx = torch.arange(grid, device=dev)[:, None] - torch.arange(grid, device=dev)[None, :] + grid - 1
        Rh = rel_h.float()[idx]; Rw = rel_w.float()[idx]
        rq_ = qf.reshape(B * H, grid, grid, hd)
        bh_ = torch.einsum("bhwc,hkc->bhwk", rq_, Rh).bfloat16().float()
        bw_ = torch.einsum("bhwc,wkc->bhwk", rq_, Rw).bfloat16().float()
        scores = (scores.reshape(B * H, grid, grid, grid, grid) + bh_[..., :, None] + bw_[..., None, :]).reshape(B * H, S, S)
        ext = 32 if grid == 14 else 64
        qext = torch.zeros(B * H, S_pad, ext, device=dev, dtype=torch.bfloat16)
        if grid == 64: rb = torch.zeros(B * H, S_pad, 64, device=dev, dtype=torch.bfloat16)
        kext = ops.make_kext(grid, dev)
        rel_hw = ops.make_rel_hw(rel_h, rel_w)
        ops.relpos_prep(q, rel_hw, bh=B * H, seq=S, seq_pad=S_pad, head_dim=hd, grid=grid, inv_scale=1.0 / scale, qext=qext, row_bias=rb)
    kvl = None
    if causal or kv_len is not None:
        i = torch.arange(S, device=dev)
        mask = torch.zeros(B, 1, S, S, device=dev, dtype=torch.bool)
        if causal: mask |= (i[None, :] > i[:, None])[None, None]
        if kv_len is not None:
            kvl = torch.tensor(kv_len, device=dev, dtype=torch.int32)
            mask |= (i[None, None, None, :] >= kvl[:, None, None, None])
        scores = scores.reshape(B, H, S, S).masked_fill(mask, float("-inf")).reshape(B * H, S, S)
    P = torch.softmax(scores, -1)
    o_ref = (P @ vf).reshape(B, H, S, hd).permute(0, 2, 1, 3).reshape(B * S, D)
    out = torch.full((B * S, D), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.attention(q, k, vt, out, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S_pad, scale=scale, causal=causal,
                  kv_len=kvl, qext=qext, kext=kext, row_bias=rb, ext_cols=ext)
    torch.cuda.synchronize()
    valid = torch.ones(B, S, dtype=torch.bool, device=dev)
    if kv_len is not None:
        valid = torch.arange(S, device=dev)[None] < kvl[:, None]
    d = (out.float() - o_ref).reshape(B, S, D)[valid]
    print(f"[{name}] attention maxerr={d.abs().max().item():.4f} mean={d.abs().mean().item():.5f} refmax={o_ref.abs().max().item():.2f} nan={torch.isnan(out.float().reshape(B,S,D)[valid]).sum().item()}", flush=True)
    if time_it:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        for _ in range(3):
            ops.attention(q, k, vt, out, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S_pad, scale=scale, causal=causal, kv_len=kvl, qext=qext, kext=kext, row_bias=rb, ext_cols=ext)
        e0.record()
        n = 10
        for _ in range(n):
            ops.attention(q, k, vt, out, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S_pad, scale=scale, causal=causal, kv_len=kvl, qext=qext, kext=kext, row_bias=rb, ext_cols=ext)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4 * B * H * S * S * hd * (0.5 if causal else 1.0)
        print(f"[{name}] attention {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TF/s (QK+PV flops only)", flush=True)

which = sys.argv[1:]
def want(n): return not which or n in which
if want("small"): run("hd64 S=128", 1, 2, 64, 128)
if want("small"): run("hd64 S=200", 1, 2, 64, 200)
if want("clip"): run("clip hd64 S=257", 2, 16, 64, 257, time_it=True)
if want("hd80"): run("hd80 S=300 noext", 1, 4, 80, 300)
if want("llama"): run("llama hd128 T=319 causal", 2, 32, 128, 319, causal=True, kv_len=[319, 250], rope=True, time_it=True)
if want("win"): run("sam window hd80 S=196", 25, 16, 80, 196, grid=14, time_it=True)
if want("glob"): run("sam global hd80 S=4096", 1, 16, 80, 4096, grid=64, time_it=True)
print("launches", _lib.launch_count())
