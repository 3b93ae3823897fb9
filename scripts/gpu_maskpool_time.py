"""Mask pooling (adjoint upsample + pool) at batch 8 x 64 proposals: time and algorithmic GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"; torch.manual_seed(0)
B, K = 8, 64
segs = torch.rand(B * K, 256, 256, device=dev).bfloat16(); emb = torch.randn(B, 4096, 256, device=dev).bfloat16()
mi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(K).contiguous()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
us = min(t(lambda: ops.maskpool(segs, emb, mi)) for _ in range(3))
nbytes = segs.numel() * 2 + emb.numel() * 2 + B * K * 256 * 2
print(f"maskpool B={B} K={K}: {us:7.1f} us  {nbytes / us / 1e3:6.0f} GB/s (algorithmic: every soft mask read once)")
