"""HBM-bound kernels: achieved GB/s (algorithmic bytes / CUDA-event time), inputs larger than L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"; torch.manual_seed(0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e-3
rows, dim = 32768 * 4, 1280      # 335 MB in + 335 MB out: larger than the 126 MB L2
x = torch.randn(rows, dim, device=dev).bfloat16(); g = torch.ones(dim, device=dev).bfloat16(); b = torch.zeros(dim, device=dev).bfloat16()
out = torch.empty_like(x)
s = t(lambda: ops.layernorm(x, g, b, 1e-6, out=out)); print(f"layernorm {rows}x{dim}: {s*1e6:.1f} us  {2*rows*dim*2/s/1e9:.0f} GB/s")
s = t(lambda: out.copy_(x)); print(f"torch copy   {rows}x{dim}: {s*1e6:.1f} us  {2*rows*dim*2/s/1e9:.0f} GB/s")
rows2 = 32768
x2 = x[:rows2]; o2 = out[:rows2]
s = t(lambda: ops.layernorm(x2, g, b, 1e-6, out=o2)); print(f"layernorm {rows2}x{dim} (L2-resident working set 168 MB): {s*1e6:.1f} us  {2*rows2*dim*2/s/1e9:.0f} GB/s")
x3 = torch.randn(2552 * 16, 4096, device=dev).bfloat16(); g3 = torch.ones(4096, device=dev).bfloat16(); o3 = torch.empty_like(x3)
s = t(lambda: ops.rmsnorm(x3, g3, 1e-6, out=o3)); print(f"rmsnorm {x3.shape[0]}x4096: {s*1e6:.1f} us  {2*x3.numel()*2/s/1e9:.0f} GB/s")
K = 64 * 8
segs = torch.rand(K, 256, 256, device=dev).bfloat16(); emb = torch.randn(8, 4096, 256, device=dev).bfloat16()
mi = torch.arange(8, device=dev, dtype=torch.int32).repeat_interleave(64).contiguous()
s = t(lambda: ops.maskpool(segs, emb, mi)); print(f"maskpool B=8 K=64: {s*1e6:.1f} us  {(K*65536*2 + 8*4096*256*2)/s/1e9:.0f} GB/s (algorithmic)")
