"""Time the SAM window attention kernel at the batch-8 shape (200 windows x 16 heads, un-partition row map as the encoder
passes it); run under LLMSEG_B200_LIB=<other build> for a same-box A/B of two library builds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
from llmseg_b200.encoders import SamEncoder
from llmseg_b200.lisa import SamCfg
dev = "cuda"; torch.manual_seed(0)
B, H, hd = 8, 16, 80; scale = hd ** -0.5
enc = SamEncoder.__new__(SamEncoder); enc.cfg, enc.device, enc._maps = SamCfg(), torch.device(dev), {}
win_map, n_win, tok2win, pad_wins = enc._window_maps(B)
S, Sp, nb = 196, 200, B * n_win
q = torch.randn(nb * H, Sp, hd, device=dev).bfloat16(); k = torch.randn_like(q); vt = torch.randn(nb * H, hd, Sp, device=dev).bfloat16()
rel = ops.make_rel_hw((torch.randn(27, hd, device=dev) * 0.1).bfloat16(), (torch.randn(27, hd, device=dev) * 0.1).bfloat16())
qext = torch.zeros(nb * H, Sp, 32, device=dev, dtype=torch.bfloat16)
ops.relpos_prep(q, rel, bh=nb * H, seq=S, seq_pad=Sp, head_dim=hd, grid=14, inv_scale=1 / scale, qext=qext)
out = torch.zeros(B * 4096, H * hd, device=dev, dtype=torch.bfloat16); kext = ops.make_kext(14, dev)
def t(fn, n=20):
    best = 1e9
    for _ in range(3):
        fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
us = t(lambda: ops.attention(q, k, vt, out, batch=nb, heads=H, head_dim=hd, seq=S, seq_pad=Sp, scale=scale, qext=qext, kext=kext,
                             ext_cols=32, out_row_map=win_map))
print(f"{os.environ.get('LLMSEG_B200_LIB', 'default lib')}: window attention {nb} windows x {H} heads: {us:7.1f} us  checksum {out.float().sum().item():.3f}")
