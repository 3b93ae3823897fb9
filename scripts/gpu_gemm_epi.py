"""Epilogue cost micro-benchmark: same GEMM shape with different fused epilogues (CUDA events, B=8 shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"; torch.manual_seed(0)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
def mk(M, N, K):
    return torch.randn(M, K, device=dev).bfloat16(), (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16(), torch.randn(N, device=dev).bfloat16()
M = 32768
a, w, b = mk(M, 5120, 1280); out = torch.empty(M, 5120, device=dev, dtype=torch.bfloat16)
for name, kw in (("none", {}), ("bias", dict(bias=b)), ("bias+relu", dict(bias=b, act="relu")), ("bias+quick_gelu", dict(bias=b, act="quick_gelu")), ("bias+gelu", dict(bias=b, act="gelu"))):
    bias = kw.pop("bias", None)
    us = t(lambda: ops.gemm(a, w, bias, out=out, **kw)); print(f"mlp1 32768x5120x1280 {name:16s} {us:7.1f} us  {2*M*5120*1280/us/1e6:6.0f} TF/s", flush=True)
a, w, b = mk(M, 1280, 5120); out = torch.empty(M, 1280, device=dev, dtype=torch.bfloat16); res = torch.randn(M, 1280, device=dev).bfloat16()
for name, kw in (("bias", dict()), ("bias+residual", dict(residual=res))):
    us = t(lambda: ops.gemm(a, w, b, out=out, **kw)); print(f"mlp2 32768x1280x5120 {name:16s} {us:7.1f} us  {2*M*5120*1280/us/1e6:6.0f} TF/s", flush=True)
Mw = 39200
a, w, b = mk(Mw, 1280, 1280); x = torch.randn(M, 1280, device=dev).bfloat16()
rmap = torch.randperm(Mw, device=dev)[:Mw].to(torch.int32); rmap = torch.where(rmap < M, rmap, torch.full_like(rmap, -1)).contiguous()
out = torch.empty(Mw, 1280, device=dev, dtype=torch.bfloat16)
for name, fn in (("bias", lambda: ops.gemm(a, w, b, out=out)), ("bias+residual", lambda: ops.gemm(a, w, b, residual=out, out=out)),
                 ("bias+res+rowmap(random)", lambda: ops.gemm(a, w, b, residual=x, out=x, out_row_map=rmap))):
    us = t(fn); print(f"proj 39200x1280x1280 {name:24s} {us:7.1f} us  {2*Mw*1280*1280/us/1e6:6.0f} TF/s", flush=True)
# QKV split cost
H, hd, S, Sp = 16, 80, 196, 200; nb = Mw // S
a, w, b = mk(Mw, 3840, 1280)
q = torch.zeros(nb * H, Sp, hd, device=dev, dtype=torch.bfloat16); k = torch.zeros_like(q); vt = torch.zeros(nb * H, hd, Sp, device=dev, dtype=torch.bfloat16)
out = torch.empty(Mw, 3840, device=dev, dtype=torch.bfloat16)
us = t(lambda: ops.gemm(a, w, b, out=out)); print(f"qkv 39200x3840x1280 plain store      {us:7.1f} us  {2*Mw*3840*1280/us/1e6:6.0f} TF/s")
us = t(lambda: ops.gemm_qkv(a, w, b, q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=Sp)); print(f"qkv 39200x3840x1280 split q/k/vT     {us:7.1f} us  {2*Mw*3840*1280/us/1e6:6.0f} TF/s")
# LLaMA batch-8 shapes (M = 8 x 319 tokens): stream-K tail on/off
Ml = 2552
for name, N, K, kw in (("o_proj", 4096, 4096, "res"), ("down", 4096, 11008, "res"), ("gate_up", 22016, 4096, "swiglu"),
                       ("qkv+rope", 12288, 4096, "qkv"), ("sam mlp2", 1280, 5120, "res32768")):
    M_ = 32768 if kw == "res32768" else Ml
    a, w, b = mk(M_, N, K)
    res = torch.randn(M_, N, device=dev).bfloat16()
    out = torch.empty(M_, N // 2 if kw == "swiglu" else N, device=dev, dtype=torch.bfloat16)
    if kw == "qkv":
        Hh = 32; q = torch.zeros(8 * Hh, 320, 128, device=dev, dtype=torch.bfloat16); k = torch.zeros_like(q)
        vt = torch.zeros(8 * Hh, 128, 320, device=dev, dtype=torch.bfloat16)
        cs = torch.randn(319, 64, device=dev).bfloat16()
        fn = lambda: ops.gemm_qkv(a, w, None, q, k, vt, heads=Hh, head_dim=128, seq_in=319, seq_pad=320, rope_cos=cs, rope_sin=cs)
    elif kw == "swiglu":
        fn = lambda: ops.gemm(a, w, None, out=out, swiglu=True)
    else:
        fn = lambda: ops.gemm(a, w, None, residual=res, out=out)
    r = [1e9, 1e9]
    for rep in range(3):   # interleaved, best of 3: the two schedules see the same clocks
        for i, use in enumerate((False, True)):
            ops.USE_GEMM_WORKSPACE = use
            r[i] = min(r[i], t(fn, 10))
    ops.USE_GEMM_WORKSPACE = True
    print(f"{name:9s} {M_}x{N}x{K}: plain {r[0]:7.1f} us ({2*M_*N*K/r[0]/1e6:5.0f} TF/s)   stream-K tail {r[1]:7.1f} us ({2*M_*N*K/r[1]/1e6:5.0f} TF/s)", flush=True)
# cost of the row-statistics epilogue (partials only / partials + in-kernel finish)
for name, M_, N, K in (("proj", 32768, 1280, 1280), ("mlp2", 32768, 1280, 5120), ("o_proj", 2552, 4096, 4096), ("down", 2552, 4096, 11008)):
    a, w, b = mk(M_, N, K)
    res = torch.randn(M_, N, device=dev).bfloat16(); out = torch.empty_like(res)
    st = ops.gemm_stats_buffer(M_, N, M_, 1e-6)
    st_p = ops.RowStats(st.t, st.parts, st.dim, st.eps, st.rms)
    r = [1e9, 1e9, 1e9]
    for rep in range(3):
        for i, so in enumerate((None, st_p, st)):
            r[i] = min(r[i], t(lambda: ops.gemm(a, w, None, residual=res, out=out, stats_out=so), 10))
    print(f"{name:7s} {M_}x{N}x{K}: no stats {r[0]:7.1f} us   partials {r[1]:7.1f} us   partials+finish {r[2]:7.1f} us", flush=True)
