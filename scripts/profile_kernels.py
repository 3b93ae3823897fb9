"""Run a few launches of one hot kernel at the bench shapes (batch 8) — target for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "attn_global"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
H, hd = 16, 80
scale = hd ** -0.5

def qkv_bufs(nb, S, S_pad):
    q = (torch.randn(nb * H, S_pad, hd, device=dev) * 1.0).bfloat16()
    k = (torch.randn(nb * H, S_pad, hd, device=dev) * 1.0).bfloat16()
    vt = (torch.randn(nb * H, hd, S_pad, device=dev) * 1.0).bfloat16()
    return q, k, vt

if which == "attn_global":
    S = 4096
    q, k, vt = qkv_bufs(B, S, S)
    rel = ops.make_rel_hw((torch.randn(127, hd, device=dev) * 0.1).bfloat16(), (torch.randn(127, hd, device=dev) * 0.1).bfloat16())
    qext = torch.zeros(B * H, S, 64, device=dev, dtype=torch.bfloat16); rb = torch.zeros_like(qext)
    ops.relpos_prep(q, rel, bh=B * H, seq=S, seq_pad=S, head_dim=hd, grid=64, inv_scale=1 / scale, qext=qext, row_bias=rb)
    out = torch.empty(B * S, H * hd, device=dev, dtype=torch.bfloat16)
    kext = ops.make_kext(64, dev)
    for _ in range(iters):
        ops.attention(q, k, vt, out, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S, scale=scale, qext=qext, kext=kext, row_bias=rb, ext_cols=64)
elif which == "attn_llama":      # LLaMA-7B causal attention at T = 319 (64-token prompts): 32 heads x 128, batch B
    Hh, hdd, S = 32, 128, 319
    Sp = (S + 7) // 8 * 8
    q = torch.randn(B * Hh, Sp, hdd, device=dev).bfloat16(); k = torch.randn_like(q); vt = torch.randn(B * Hh, hdd, Sp, device=dev).bfloat16()
    out = torch.empty(B * S, Hh * hdd, device=dev, dtype=torch.bfloat16)
    for _ in range(iters):
        ops.attention(q, k, vt, out, batch=B, heads=Hh, head_dim=hdd, seq=S, seq_pad=Sp, scale=hdd ** -0.5, causal=True)
elif which == "attn_window":
    S, S_pad, nb = 196, 200, 25 * B
    q, k, vt = qkv_bufs(nb, S, S_pad)
    rel = ops.make_rel_hw((torch.randn(27, hd, device=dev) * 0.1).bfloat16(), (torch.randn(27, hd, device=dev) * 0.1).bfloat16())
    qext = torch.zeros(nb * H, S_pad, 32, device=dev, dtype=torch.bfloat16)
    ops.relpos_prep(q, rel, bh=nb * H, seq=S, seq_pad=S_pad, head_dim=hd, grid=14, inv_scale=1 / scale, qext=qext)
    out = torch.empty(nb * S, H * hd, device=dev, dtype=torch.bfloat16)
    kext = ops.make_kext(14, dev)
    for _ in range(iters):
        ops.attention(q, k, vt, out, batch=nb, heads=H, head_dim=hd, seq=S, seq_pad=S_pad, scale=scale, qext=qext, kext=kext, ext_cols=32)
elif which == "gemm_qkv_sam":    # SAM windowed layer's QKV projection as the encoder runs it: folded norm + bias + window scatter
    from llmseg_b200.encoders import SamEncoder
    S, ws = 4096, 14
    x = torch.randn(B * S, 1280, device=dev).bfloat16(); w = (torch.randn(3840, 1280, device=dev) / 1280 ** 0.5).bfloat16()
    bias = torch.randn(3840, device=dev).bfloat16()
    enc = SamEncoder.__new__(SamEncoder)
    from llmseg_b200.lisa import SamCfg
    enc.cfg, enc.device, enc._maps = SamCfg(), torch.device(dev), {}
    win_map, n_win, tok2win, pad_wins = enc._window_maps(B)
    nb, sw, sw_pad = B * n_win, ws * ws, 200
    q = torch.zeros(nb * H, sw_pad, hd, device=dev, dtype=torch.bfloat16); k = torch.zeros_like(q)
    vt = torch.zeros(nb * H, hd, sw_pad, device=dev, dtype=torch.bfloat16)
    st = ops.norm_stats(x, 1e-6)
    for _ in range(iters):
        ops.gemm_qkv(x, w, bias, q, k, vt, heads=H, head_dim=hd, seq_in=sw, seq_pad=sw_pad, row_map=tok2win, row_stats=st)
elif which.startswith("gemm"):
    shapes = {"gemm_up1": (1048576, 256, 256), "gemm_qkv": (4096 * B, 3840, 1280), "gemm_mlp1": (4096 * B, 5120, 1280), "gemm_mlp2": (4096 * B, 1280, 5120),
              "gemm_llama_gu": (319 * B, 22016, 4096), "gemm_llama_down": (319 * B, 4096, 11008),
              "gemm_proj": (4096 * B, 1280, 1280)}
    M, N, K = shapes[which]
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    for _ in range(iters):
        if which == "gemm_mlp1": ops.gemm(a, w, bias, act="gelu")
        elif which == "gemm_llama_gu": ops.gemm(a, w, None, swiglu=True)
        elif which == "gemm_proj":   # SAM attention out-projection as the encoder runs it: in-place residual + row statistics
            if _ == 0:
                x = torch.randn(M, N, device=dev).bfloat16(); st = ops.gemm_stats_buffer(M, N, M, 1e-6)
            ops.gemm(a, w, bias, residual=x, out=x, stats_out=st)
        else: ops.gemm(a, w, bias)
elif which == "amg_tok2img":     # mask decoder token->image attention, 256 prompts, per-prompt K / V as column views
    P = 256
    q = torch.randn(P * 7, 128, device=dev).bfloat16(); kv = torch.randn(P * 4096, 256, device=dev).bfloat16()
    for _ in range(iters): ops.tok2img_attention(q, kv[:, :128], kv[:, 128:], P, False)
elif which == "amg_img2tok":
    P = 256
    qb = torch.randn(P * 4096, 384, device=dev).bfloat16(); k = torch.randn(P * 7, 128, device=dev).bfloat16(); v = torch.randn_like(k)
    for _ in range(iters): ops.img2tok_attention(qb[:, 256:384], k, v, P, False)
elif which == "amg_upscale":     # fused LayerNorm2d+GELU -> ConvTranspose #2 -> hyper-network product, 256 prompts
    P = 256
    u1 = torch.randn(P * 4096, 256, device=dev).bfloat16(); gm = torch.ones(64, device=dev).bfloat16(); bt = torch.zeros(64, device=dev).bfloat16()
    w2 = (torch.randn(128, 64, device=dev) / 8).bfloat16(); b2 = torch.zeros(128, device=dev).bfloat16(); hy = torch.randn(P, 4, 32, device=dev).bfloat16()
    for _ in range(iters): ops.upscale_logits(u1, gm, bt, w2, b2, hy, P)
elif which == "maskpool":
    K = 64
    segs = torch.rand(B * K, 256, 256, device=dev).bfloat16(); emb = torch.randn(B, 4096, 256, device=dev).bfloat16()
    mi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(K).contiguous()
    for _ in range(iters): ops.maskpool(segs, emb, mi)
elif which == "layernorm":
    x = torch.randn(4096 * B, 1280, device=dev).bfloat16(); g = torch.ones(1280, device=dev).bfloat16(); b = torch.zeros(1280, device=dev).bfloat16()
    for _ in range(iters): ops.layernorm(x, g, b, 1e-6)
torch.cuda.synchronize()
print("done", which)
