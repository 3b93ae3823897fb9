"""Achieved HBM GB/s of the norm kernels (algorithmic bytes: read + write of the rows) — register-resident kernel vs
the bulk-copy-staged streaming kernel (LLMSEG_NORM_STREAM=0/1, interleaved), next to a torch copy of the same bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"; torch.manual_seed(0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
cases = [("layernorm", 131072, 1280), ("layernorm", 32768, 1280), ("layernorm", 65536, 1024), ("layernorm", 262144, 256),
         ("rmsnorm", 40832, 4096), ("rmsnorm", 8192, 4096), ("norm_stats", 131072, 1280), ("norm_stats rms", 40832, 4096)]
for name, rows, dim in cases:
    x = torch.randn(rows, dim, device=dev).bfloat16(); g = torch.ones(dim, device=dev).bfloat16(); b = torch.zeros(dim, device=dev).bfloat16()
    out = torch.empty_like(x)
    if name == "layernorm": fn = lambda: ops.layernorm(x, g, b, 1e-6, out=out); nbytes = 2 * x.numel() * 2
    elif name == "rmsnorm": fn = lambda: ops.rmsnorm(x, g, 1e-6, out=out); nbytes = 2 * x.numel() * 2
    else:
        st = torch.empty(rows, 2, device=dev); rms = "rms" in name
        fn = lambda: ops.norm_stats(x, 1e-6, rms=rms, out=st); nbytes = x.numel() * 2
    best = {"0": 1e9, "1": 1e9}
    for r in range(3):
        for v in ("0", "1"):
            os.environ["LLMSEG_NORM_STREAM"] = v
            best[v] = min(best[v], t(fn))
    cp = t(lambda: out.copy_(x)) if "stats" not in name else None
    print(f"{name:15s} {rows}x{dim}: registers {best['0']:7.1f} us {nbytes/best['0']/1e3:6.0f} GB/s   streamed {best['1']:7.1f} us {nbytes/best['1']/1e3:6.0f} GB/s"
          + (f"   torch copy {cp:7.1f} us {nbytes/cp/1e3:6.0f} GB/s" if cp else ""), flush=True)
