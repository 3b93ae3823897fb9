"""Throughput of the plain GEMM (no epilogue extras) over a list of MxNxK shapes, stream-K tail off/on, interleaved."""
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
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(2552, 4096, 4096), (2560, 4096, 4096), (18944, 4096, 4096), (2552, 4096, 11008), (2552, 12288, 4096)]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, None, out=out)
    best = {"0": 1e9, "1": 1e9}
    for r in range(4):
        for v in ("0", "1"):
            os.environ["LLMSEG_GEMM_STREAMK"] = v
            best[v] = min(best[v], t(fn))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: streamk=0 {best['0']:7.1f} us ({fl/best['0']/1e6:5.0f} TF/s)   =1 {best['1']:7.1f} us ({fl/best['1']/1e6:5.0f} TF/s)", flush=True)
