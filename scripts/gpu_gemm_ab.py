"""Interleaved A/B of one GEMM option (env var read per launch) on the encoder shapes as the encoders run them:
in-place residual + epilogue row statistics finished in-kernel.  usage: gpu_gemm_ab.py ENV_NAME [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"; torch.manual_seed(0)
name = sys.argv[1] if len(sys.argv) > 1 else "LLMSEG_GEMM_TMA_RES"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3

cases = (("sam proj", 32768, 1280, 1280, False), ("sam mlp2", 32768, 1280, 5120, False), ("clip out", 2056, 1024, 1024, False),
         ("clip fc2", 2056, 1024, 4096, False), ("llama o_proj", 2552, 4096, 4096, True), ("llama down", 2552, 4096, 11008, True),
         ("dino proj", 32776, 1024, 1024, False), ("dino fc2", 32776, 1024, 4096, False))
# lin1-type: folded norm + bias + GELU (no residual)
for label, M, N, K in (("sam lin1", 32768, 5120, 1280), ("dino fc1", 32776, 4096, 1024)):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    stn = ops.norm_stats(a, 1e-6)
    fn = lambda: ops.gemm(a, w, b, act="gelu", row_stats=stn)
    best = {"0": 1e9, "1": 1e9}
    for r in range(rounds):
        for v in ("0", "1"):
            os.environ[name] = v
            best[v] = min(best[v], t(fn))
    os.environ.pop(name, None)
    fl = 2.0 * M * N * K
    print(f"{label:13s} {M}x{N}x{K}: {name}=0 {best['0']:7.1f} us ({fl/best['0']/1e6:5.0f} TF/s)   =1 {best['1']:7.1f} us ({fl/best['1']/1e6:5.0f} TF/s)   {100*(best['0']/best['1']-1):+5.1f} %", flush=True)
# residual, no bias, no statistics: LLaMA o_proj / down_proj with the text branch's norms un-folded (epilogue variant 3)
for label, M, N, K in (("llama o_proj*", 2552, 4096, 4096), ("llama down*", 2552, 4096, 11008)):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    x = torch.randn(M, N, device=dev).bfloat16()
    fn = lambda: ops.gemm(a, w, None, residual=x, out=x)
    best = {"0": 1e9, "1": 1e9}
    for r in range(rounds):
        for v in ("0", "1"):
            os.environ[name] = v
            best[v] = min(best[v], t(fn))
    os.environ.pop(name, None)
    fl = 2.0 * M * N * K
    print(f"{label:13s} {M}x{N}x{K}: {name}=0 {best['0']:7.1f} us ({fl/best['0']/1e6:5.0f} TF/s)   =1 {best['1']:7.1f} us ({fl/best['1']/1e6:5.0f} TF/s)   {100*(best['0']/best['1']-1):+5.1f} %", flush=True)
for label, M, N, K, rms in cases:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = None if rms else torch.randn(N, device=dev).bfloat16()
    x = torch.randn(M, N, device=dev).bfloat16()
    st = ops.gemm_stats_buffer(M, N, M, 1e-6, rms=rms)
    fn = lambda: ops.gemm(a, w, b, residual=x, out=x, stats_out=st)
    best = {"0": 1e9, "1": 1e9}
    for r in range(rounds):
        for v in ("0", "1"):
            os.environ[name] = v
            best[v] = min(best[v], t(fn))
    os.environ.pop(name, None)
    fl = 2.0 * M * N * K
    print(f"{label:13s} {M}x{N}x{K}: {name}=0 {best['0']:7.1f} us ({fl/best['0']/1e6:5.0f} TF/s)   =1 {best['1']:7.1f} us ({fl/best['1']/1e6:5.0f} TF/s)   {100*(best['0']/best['1']-1):+5.1f} %", flush=True)
