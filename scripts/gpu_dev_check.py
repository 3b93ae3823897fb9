"""Developer GPU check (run under gpurun): tcgen05 layout probes, GEMM / norm numerics and timing."""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops, _lib

torch.manual_seed(0)
dev = "cuda"
print(torch.cuda.get_device_name(0), flush=True)

def probe():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/probe/libprobe.so")
    if not os.path.exists(path):
        print("no probe lib"); return
    pl = C.CDLL(path)
    pl.probe_mma.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    for mode in (0, 1):
        for sw in (128, 64, 32):
            KA = sw // 2
            for N in (64, 208, 256):
                A = torch.randn(128, KA, device=dev).bfloat16()
                B = torch.randn(N, KA, device=dev).bfloat16()
                D = torch.full((128, N), float("nan"), device=dev)
                rc = pl.probe_mma(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, KA, sw, mode, None)
                torch.cuda.synchronize()
                ref = A.float() @ B.float().T
                err = (D - ref).abs().max().item()
                print(f"probe mode={mode} sw={sw} N={N} rc={rc} maxerr={err:.3e}", flush=True)

def gemm_check(M, N, K, bias=True, act=None, residual=False, swiglu=False):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16() if bias else None
    n_out = N // 2 if swiglu else N
    r = torch.randn(M, n_out, device=dev).bfloat16() if residual else None
    out = ops.gemm(a, w, b, act=act, residual=r, swiglu=swiglu)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    if b is not None: ref = ref + b.float()
    ref = ref.bfloat16().float()
    if swiglu:
        g, u = ref[:, 0::2], ref[:, 1::2]
        ref = (torch.nn.functional.silu(g).bfloat16().float() * u)
    if act == "gelu": ref = torch.nn.functional.gelu(ref).bfloat16().float()
    if act == "quick_gelu": ref = (ref * torch.sigmoid(1.702 * ref)).bfloat16().float()
    if act == "relu": ref = torch.relu(ref)
    if r is not None: ref = ref + r.float()
    err = (out.float() - ref).abs().max().item()
    print(f"gemm M={M} N={N} K={K} bias={bias} act={act} res={residual} swiglu={swiglu} maxerr={err:.4f} refmax={ref.abs().max().item():.2f}", flush=True)
    return err

def gemm_time(M, N, K, iters=20):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): ops.gemm(a, w, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2 * M * N * K / ms / 1e9
    # cuBLAS comparison
    for _ in range(3): torch.matmul(a, w.T)
    e0.record()
    for _ in range(iters): torch.matmul(a, w.T)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"gemm time M={M} N={N} K={K}: {ms*1e3:.1f} us = {tf:.0f} TF/s   (cuBLAS {ms2*1e3:.1f} us = {2*M*N*K/ms2/1e9:.0f} TF/s)", flush=True)

def norm_check():
    for rows, dim, eps in ((4096, 1280, 1e-6), (257, 1024, 1e-5), (319, 4096, 1e-6), (64, 256, 1e-5)):
        x = (torch.randn(rows, dim, device=dev) * 2 + 0.5).bfloat16()
        g = torch.randn(dim, device=dev).bfloat16(); b = torch.randn(dim, device=dev).bfloat16()
        y = ops.layernorm(x, g, b, eps)
        ref = torch.nn.functional.layer_norm(x.float(), (dim,), g.float(), b.float(), eps)
        print(f"layernorm {rows}x{dim} maxerr={(y.float()-ref).abs().max().item():.4f}")
        y = ops.rmsnorm(x, g, eps)
        xf = x.float(); ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).bfloat16().float() * g.float()
        print(f"rmsnorm {rows}x{dim} maxerr={(y.float()-ref).abs().max().item():.4f}", flush=True)

which = sys.argv[1:] or ["probe", "gemm", "norm", "time"]
if "probe" in which: probe()
if "gemm" in which:
    gemm_check(128, 128, 64, bias=False)
    gemm_check(128, 256, 128, bias=False)
    gemm_check(256, 512, 1280)
    gemm_check(4096, 3840, 1280)
    gemm_check(300, 1032, 1288, act="gelu", residual=True)
    gemm_check(257, 1024, 1024, act="quick_gelu")
    gemm_check(319, 22016, 4096, bias=False, swiglu=True)
    gemm_check(4900, 1280, 1280, residual=True)
    gemm_check(64, 256, 256, act="relu")
if "norm" in which: norm_check()
if "time" in which:
    gemm_time(4096, 3840, 1280); gemm_time(4096, 5120, 1280); gemm_time(4096, 1280, 5120)
    gemm_time(8192, 8192, 8192); gemm_time(319, 12288, 4096); gemm_time(319, 4096, 11008)
    gemm_time(32768, 5120, 1280)
print("launches", _lib.launch_count())
