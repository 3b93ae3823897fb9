"""Time the SAM windowed layer's QKV projection as the encoder runs it (folded norm + bias + window row map) and the
LLaMA QKV + RoPE projection; run under LLMSEG_B200_LIB=<other build> for a same-box A/B of two library builds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
from llmseg_b200.encoders import SamEncoder
from llmseg_b200.lisa import SamCfg
dev = "cuda"; torch.manual_seed(0)
B, H, hd, S, ws = 8, 16, 80, 4096, 14
x = torch.randn(B * S, 1280, device=dev).bfloat16(); w = (torch.randn(3840, 1280, device=dev) / 1280 ** 0.5).bfloat16()
bias = torch.randn(3840, device=dev).bfloat16()
enc = SamEncoder.__new__(SamEncoder); enc.cfg, enc.device, enc._maps = SamCfg(), torch.device(dev), {}
win_map, n_win, tok2win, pad_wins = enc._window_maps(B)
nb, sw, sw_pad = B * n_win, ws * ws, 200
q = torch.zeros(nb * H, sw_pad, hd, device=dev, dtype=torch.bfloat16); k = torch.zeros_like(q)
vt = torch.zeros(nb * H, hd, sw_pad, device=dev, dtype=torch.bfloat16)
st = ops.norm_stats(x, 1e-6)
def t(fn, n=20):
    best = 1e9
    for _ in range(3):
        fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
us = t(lambda: ops.gemm_qkv(x, w, bias, q, k, vt, heads=H, head_dim=hd, seq_in=sw, seq_pad=sw_pad, row_map=tok2win, row_stats=st))
print(f"{os.environ.get('LLMSEG_B200_LIB', 'default lib')}: SAM window QKV 32768x3840x1280: {us:7.1f} us  {2 * 32768 * 3840 * 1280 / us / 1e6:6.0f} TF/s  checksum {q.float().sum().item():.3f} {vt.float().sum().item():.3f}")
