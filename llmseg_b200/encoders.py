"""Host-side orchestration of the three encoders on top of the sm_100a kernels (ops.py).

Each class takes a state dict with the reference's parameter names, re-lays the weights out once
for the kernels (bf16, fused QKV / interleaved gate-up, conv weights flattened for conv-as-GEMM),
and exposes a forward that only enqueues llmseg_* kernels.

  SamEncoder   reference model/segment_anything/modeling/image_encoder.py:110-125 (ViT-H/16 @1024)
  Dinov2Encoder reference model/LISA.py:186-199,244-245 (hub dinov2_vitl14 @896 + lisa_dino_conv)
  ClipTower    reference model/llava/model/multimodal_encoder/clip_encoder.py:31-60 + mm_projector
               (llava_arch.py:93-96); arithmetic = transformers CLIPVisionTransformer
  LlamaDecoder reference model/llava/model/language_model/llava_llama.py:93-102; arithmetic =
               transformers LlamaModel (lm_head is skipped: dead work at inference, LISA.py:283,318)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops

import os

Tensor = torch.Tensor
BF16 = torch.bfloat16
# LayerNorm / RMSNorm in front of a projection can be folded into that GEMM (ops.fold_norm + ops.norm_stats):
# the norm pass over the residual stream becomes a statistics-only read.  LLMSEG_FOLD_NORM selects where:
#   "sam" (default)  SAM ViT-H only    "image"  SAM + DINOv2    "all" / "1"  every encoder    "0"  nowhere (the reference's literal op order)
# Why not everywhere: folding moves a bf16 rounding from the normalised ACTIVATIONS (independent per token, averaged
# away by attention) to the gamma-scaled WEIGHTS (the same perturbation for every token, so it adds up coherently).
# tests/parity_bisect.py (profiles/round2_parity_bisect.md): with the text branch folded, its contribution to the
# |pred_similarity - fp32 oracle| error is 3x larger (1.2e-3 vs 0.4e-3 mean) — the level of the reference's own bf16
# path — while the image branch's contribution is negligible either way (2e-4), and the image branch is where the
# norm passes cost time (64 x [32768, 1280] at batch 8 vs 64 x [2552, 4096]).
# DINOv2 (variant B) feeds the selector through ONE folded GEMM (final norm + lisa_dino_conv) after 24 LayerScale'd
# blocks, and there folding does show at the outputs: full depth, similarity 2.9e-3 folded vs 1.5e-3 un-folded against
# the fp32 oracle (reference bf16 path: 1.2e-3; profiles/round2_parity_bisect.md) — so "sam" (the default) folds the SAM
# encoder only; "image" adds DINOv2, "all" / "1" every encoder, "0" none.
_FOLD = os.environ.get("LLMSEG_FOLD_NORM", "sam").lower()
FOLD_NORM_SAM = _FOLD in ("sam", "image", "all", "1")
FOLD_NORM_IMAGE = _FOLD in ("image", "all", "1")
FOLD_NORM_TEXT = _FOLD in ("all", "1")
# Row-restricted tail of the last LLaMA layer (exact: row-wise ops commute with the [SEG] gather).  Off by default:
# measured no gain at batch 8 (80.30 vs 80.33 ms/step, profiles/r02b) — the M=8 GEMMs stream the same 300 MB of
# weights — and it moves bf16 rounding points (the statistics come from the bf16 rows instead of the fp32 epilogue).
LAST_LAYER_ROWS = os.environ.get("LLMSEG_LAST_LAYER_ROWS", "0") != "0"
# SAM windowed layers: private k / vT buffers per layer with the constant padding rows written once (see SamEncoder).
# The buffers (5.7 GB at batch 8) belong to the CALLER's plan (`private` argument of SamEncoder.forward): they are freed
# when that plan — and the CUDA graph that bakes their addresses in — is evicted, instead of accumulating per batch size.
KV_PREFILL = os.environ.get("LLMSEG_KV_PREFILL", "1") != "0"
KV_PREFILL_MAX_BYTES = 8 << 30


def _dev(t: Tensor, device) -> Tensor:
    return t.detach().to(device=device, dtype=BF16).contiguous()


class _Scratch:
    """Zero-initialised attention staging buffers, cached by shape.  The pad rows/columns of
    q/k/vt/qext are never written by the kernels and must stay zero (finite) for the masked tiles."""

    def __init__(self, device):
        self.device = device
        self._bufs: Dict[tuple, Tensor] = {}

    def zeros(self, tag: str, *shape) -> Tensor:
        key = (tag,) + tuple(shape)
        buf = self._bufs.get(key)
        if buf is None:
            buf = torch.zeros(shape, dtype=BF16, device=self.device)
            self._bufs[key] = buf
        return buf

    def stats(self, rows: int, parts: int = 1, tag: str = "a") -> Tensor:
        """fp32 row-statistics buffer: [rows, 2] for ops.norm_stats, [rows, parts, 2] for GEMM-epilogue partials."""
        key = ("stats", tag, rows, parts)
        buf = self._bufs.get(key)
        if buf is None:
            buf = torch.empty((rows, parts, 2), dtype=torch.float32, device=self.device)
            self._bufs[key] = buf
        return buf

    def gemm_stats(self, M: int, N: int, eps: float, tag: str, rms: bool = False):
        """ops.RowStats buffer for gemm(stats_out=...) of an [M, *] x [N, *] problem (cached)."""
        parts = ops.gemm_stats_parts(M, N)
        return ops.gemm_stats_buffer(M, N, M, eps, rms=rms, out=self.stats(M, parts + 1, tag))


# ==============================================================================================
# SAM ViT image encoder
# ==============================================================================================
class SamEncoder:
    def __init__(self, sd: Dict[str, Tensor], cfg, device, prefix: str = ""):
        self.cfg, self.device = cfg, device
        self.fold = FOLD_NORM_SAM
        D = cfg.embed_dim
        self.heads, self.hd = cfg.num_heads, cfg.embed_dim // cfg.num_heads
        if self.hd != 80:
            raise ValueError(f"SAM attention kernel is instantiated for head_dim 80 (ViT-H), got {self.hd}")
        g = cfg.grid
        if g != 64 or cfg.window_size != 14:
            raise ValueError("rel-pos score extension supports the 64x64 grid with 14x14 windows (SAM @1024)")
        p = cfg.patch_size
        d = lambda k: _dev(sd[prefix + k], device)
        self.w_patch = d("patch_embed.proj.weight").reshape(D, 3 * p * p).contiguous()
        self.b_patch = d("patch_embed.proj.bias")
        self.pos = d("pos_embed").reshape(g * g, D).contiguous()
        self.blocks = []
        for i in range(cfg.depth):
            bp = f"blocks.{i}."
            blk = dict(
                window=0 if i in cfg.global_attn_indexes else cfg.window_size,
                ln1_w=d(bp + "norm1.weight"), ln1_b=d(bp + "norm1.bias"),
                ln2_w=d(bp + "norm2.weight"), ln2_b=d(bp + "norm2.bias"),
                w_qkv=d(bp + "attn.qkv.weight"), b_qkv=d(bp + "attn.qkv.bias"),
                w_proj=d(bp + "attn.proj.weight"), b_proj=d(bp + "attn.proj.bias"),
                rel_hw=ops.make_rel_hw(d(bp + "attn.rel_pos_h"), d(bp + "attn.rel_pos_w")),
                w1=d(bp + "mlp.lin1.weight"), b1=d(bp + "mlp.lin1.bias"),
                w2=d(bp + "mlp.lin2.weight"), b2=d(bp + "mlp.lin2.bias"),
            )
            if self.fold:
                blk["f_qkv"] = ops.fold_norm(blk["w_qkv"], blk["ln1_w"], blk["ln1_b"], blk["b_qkv"])
                blk["f_1"] = ops.fold_norm(blk["w1"], blk["ln2_w"], blk["ln2_b"], blk["b1"])
                # the padding keys/values of a window are the projection of a zero token = the bias, in bf16
                blk["b_qkv_pad"] = blk["b_qkv"]
                del blk["w_qkv"], blk["w1"]
            self.blocks.append(blk)
        self.w_neck1 = d("neck.0.weight").reshape(cfg.out_chans, D).contiguous()
        self.ln_n1 = (d("neck.1.weight"), d("neck.1.bias"))
        # conv3x3 weight [out, in, ky, kx] -> [out, (ky, kx, in)] to match im2col3x3's column order
        self.w_neck2 = d("neck.2.weight").permute(0, 2, 3, 1).reshape(cfg.out_chans, 9 * cfg.out_chans).contiguous()
        self.ln_n2 = (d("neck.3.weight"), d("neck.3.bias"))
        self.kext_win = ops.make_kext(14, device)
        self.kext_glb = ops.make_kext(64, device)
        self.scratch = _Scratch(device)
        self._maps: Dict[int, tuple] = {}

    def _window_maps(self, B: int):
        """win_src[r]: image token feeding window row r (or -1 = zero padding token);
        unwin_dst[r]: token row receiving window row r (or -1 = cropped).  reference
        image_encoder.py:263-318 as index maps."""
        if B not in self._maps:
            g, ws = self.cfg.grid, self.cfg.window_size
            nw = (g + ws - 1) // ws
            b = torch.arange(B).view(B, 1, 1, 1, 1)
            wy = torch.arange(nw).view(1, nw, 1, 1, 1)
            wx = torch.arange(nw).view(1, 1, nw, 1, 1)
            ty = torch.arange(ws).view(1, 1, 1, ws, 1)
            tx = torch.arange(ws).view(1, 1, 1, 1, ws)
            y, x = wy * ws + ty, wx * ws + tx
            tok = b * (g * g) + y * g + x
            tok = torch.where((y < g) & (x < g), tok, torch.full_like(tok, -1))
            m = tok.reshape(-1).to(torch.int32)
            valid = m >= 0
            pos = torch.arange(m.numel(), dtype=torch.int32)
            tok2win = torch.empty(B * g * g, dtype=torch.int32)
            tok2win[m[valid].long()] = pos[valid]          # token row -> (window, slot) position
            pad_wins = (~valid).view(B * nw * nw, ws * ws).any(dim=1).nonzero().flatten().to(torch.int32)
            dev = self.device
            self._maps[B] = (m.to(dev), nw * nw, tok2win.to(dev), pad_wins.to(dev))
        return self._maps[B]

    def forward(self, images: Tensor, private: Optional[dict] = None) -> Tensor:
        """[B,3,1024,1024] bf16 -> token-major neck output [B, 4096, out_chans] bf16 (NHWC).
        private: a dict owned by the caller's per-batch-size plan; with it (and LLMSEG_KV_PREFILL) every windowed
        layer keeps its own k / vT buffers there, padding rows written once."""
        cfg = self.cfg
        B = images.shape[0]
        g, D, H, hd = cfg.grid, cfg.embed_dim, self.heads, self.hd
        S = g * g
        a = ops.patchify(images.contiguous(), cfg.patch_size, 3 * cfg.patch_size ** 2)
        # stA / stB: per-row (sum, sum of squares) partials of the residual stream, written by the epilogue of
        # whichever GEMM produced it (patch embed / lin2 -> stA for norm1, proj -> stB for norm2)
        stA = self.scratch.gemm_stats(B * S, D, cfg.ln_eps, "a") if self.fold else None
        stB = self.scratch.gemm_stats(B * S, D, cfg.ln_eps, "b") if self.fold else None
        x = ops.gemm(a, self.w_patch, self.b_patch, residual=self.pos, res_mod=S, stats_out=stA)
        del a
        win_map, n_win, tok2win, pad_wins = self._window_maps(B)
        scale = hd ** -0.5
        for li, blk in enumerate(self.blocks):
            if blk["window"] > 0:
                ws = blk["window"]
                sw, sw_pad = ws * ws, (ws * ws + 7) // 8 * 8
                nb = B * n_win
                # The 64x64 grid pads to 70x70 for 14x14 windows (image_encoder.py:263-288): 18% of the
                # window rows are zero tokens.  They never enter a GEMM here: LN1 and the QKV projection run
                # on the 4096 real tokens per image and scatter into (window, slot) order; the padding keys /
                # values equal the projection bias exactly (zero input), their queries are cropped again
                # (image_encoder.py:291-318), and attention writes straight back in token order.
                q = self.scratch.zeros("q", nb * H, sw_pad, hd)
                # The padding keys/values of a layer are constants of its weights (the projection bias).  With
                # KV_PREFILL every windowed layer owns its k / vT buffers (2 x 102 MB per layer at batch 8, 5.7 GB
                # for the 28 layers — HBM is not the scarce resource here), the padding rows are written once and
                # the per-forward fill_kv_rows launch (35 us x 28) disappears from the step.
                prefill = (private is not None and KV_PREFILL and
                           2 * nb * H * sw_pad * hd * 2 * len(self.blocks) <= KV_PREFILL_MAX_BYTES)
                if prefill:
                    if ("k", li) not in private:
                        private[("k", li)] = torch.zeros((nb * H, sw_pad, hd), dtype=BF16, device=self.device)
                        private[("vt", li)] = torch.zeros((nb * H, hd, sw_pad), dtype=BF16, device=self.device)
                    k, vt = private[("k", li)], private[("vt", li)]
                else:
                    k = self.scratch.zeros("k", nb * H, sw_pad, hd)
                    vt = self.scratch.zeros("vt", nb * H, hd, sw_pad)
                qext = self.scratch.zeros("qext_w", nb * H, sw_pad, 32)
                if self.fold:
                    wq, bq = blk["f_qkv"]
                    ops.gemm_qkv(x, wq, bq, q, k, vt, heads=H, head_dim=hd, seq_in=sw, seq_pad=sw_pad,
                                 row_map=tok2win, row_stats=stA)
                    o = self.scratch.zeros("o", B * S, D)
                else:
                    h = ops.layernorm(x, blk["ln1_w"], blk["ln1_b"], cfg.ln_eps)
                    ops.gemm_qkv(h, blk["w_qkv"], blk["b_qkv"], q, k, vt, heads=H, head_dim=hd, seq_in=sw,
                                 seq_pad=sw_pad, row_map=tok2win)
                    o = h  # reuse the LN output buffer for the attention output (same shape)
                if not (prefill and ("filled", li) in private):
                    ops.fill_kv_rows(k, vt, blk["b_qkv"], win_map, batch=nb, heads=H, head_dim=hd, seq_in=sw,
                                     seq_pad=sw_pad, seq_ids=pad_wins)
                    if prefill and not torch.cuda.is_current_stream_capturing():
                        private[("filled", li)] = True
                ops.relpos_prep(q, blk["rel_hw"], bh=nb * H, seq=sw, seq_pad=sw_pad, head_dim=hd, grid=ws,
                                inv_scale=1.0 / scale, qext=qext)
                ops.attention(q, k, vt, o, batch=nb, heads=H, head_dim=hd, seq=sw, seq_pad=sw_pad, scale=scale,
                              qext=qext, kext=self.kext_win, ext_cols=32, out_row_map=win_map)
                ops.gemm(o, blk["w_proj"], blk["b_proj"], residual=x, out=x, stats_out=stB)
            else:
                q = self.scratch.zeros("qg", B * H, S, hd)
                k = self.scratch.zeros("kg", B * H, S, hd)
                vt = self.scratch.zeros("vtg", B * H, hd, S)
                qext = self.scratch.zeros("qext_g", B * H, S, 64)
                rb = self.scratch.zeros("rb_g", B * H, S, 64)
                if self.fold:
                    wq, bq = blk["f_qkv"]
                    ops.gemm_qkv(x, wq, bq, q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=S, row_stats=stA)
                    o = self.scratch.zeros("o", B * S, D)
                else:
                    h = ops.layernorm(x, blk["ln1_w"], blk["ln1_b"], cfg.ln_eps)
                    ops.gemm_qkv(h, blk["w_qkv"], blk["b_qkv"], q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=S)
                    o = h
                ops.relpos_prep(q, blk["rel_hw"], bh=B * H, seq=S, seq_pad=S, head_dim=hd, grid=g,
                                inv_scale=1.0 / scale, qext=qext, row_bias=rb)
                ops.attention(q, k, vt, o, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S, scale=scale,
                              qext=qext, kext=self.kext_glb, row_bias=rb, ext_cols=64)
                ops.gemm(o, blk["w_proj"], blk["b_proj"], residual=x, out=x, stats_out=stB)
            if self.fold:
                w1, b1 = blk["f_1"]
                m = ops.gemm(x, w1, b1, act="gelu", row_stats=stB)
            else:
                h = ops.layernorm(x, blk["ln2_w"], blk["ln2_b"], cfg.ln_eps)
                m = ops.gemm(h, blk["w1"], blk["b1"], act="gelu")
                del h
            ops.gemm(m, blk["w2"], blk["b2"], residual=x, out=x, stats_out=stA)
            del m
        y = ops.gemm(x, self.w_neck1)
        y = ops.layernorm(y, self.ln_n1[0], self.ln_n1[1], 1e-6)
        y = ops.im2col3x3(y, B, g, g)
        y = ops.gemm(y, self.w_neck2)
        y = ops.layernorm(y, self.ln_n2[0], self.ln_n2[1], 1e-6)
        return y.view(B, S, cfg.out_chans)


# ==============================================================================================
# DINOv2 ViT-L/14 image encoder + lisa_dino_conv ("variant B": the branch the checked-in
# reference model_forward takes, model/LISA.py:186-199,244-245)
# ==============================================================================================
def _resample_pos_embed(pos: Tensor, grid: int, offset: float) -> Tensor:
    """Bicubic resampling of DINOv2's stored position table to a grid x grid image
    (DinoVisionTransformer.interpolate_pos_encoding; `offset` = its interpolate_offset, 0.1 on the hub
    default).  A one-off weight transform at load, in fp32."""
    n = pos.shape[1] - 1
    m = int(round(math.sqrt(n)))
    if m * m != n:
        raise ValueError(f"pos_embed holds {n} patch positions, not a square grid")
    pos = pos.float()
    if m == grid:
        return pos[0]
    D = pos.shape[-1]
    tbl = pos[:, 1:].reshape(1, m, m, D).permute(0, 3, 1, 2)
    kw = dict(scale_factor=(float(grid + offset) / m,) * 2) if offset else dict(size=(grid, grid))
    tbl = torch.nn.functional.interpolate(tbl, mode="bicubic", align_corners=False, **kw)
    if tbl.shape[-2:] != (grid, grid):
        raise ValueError(f"pos_embed resampling produced {tuple(tbl.shape[-2:])}, wanted {grid}x{grid}")
    return torch.cat([pos[0, :1], tbl.permute(0, 2, 3, 1).reshape(grid * grid, D)], dim=0)


class Dinov2Encoder:
    """hub `dinov2_vitl14.forward_features(...)['x_norm_patchtokens']` followed by the 1x1
    `lisa_dino_conv`, on the same kernels as the CLIP tower (head_dim 64, no rel-pos):

      * patch conv bias and cls_token are folded into the (resampled) position table, which the patch
        GEMM adds as a per-token residual
      * LayerScale is folded into the weights of the projection it scales (`ls*(W a + b)`)
      * norm1/norm2 and the final norm are folded into the consuming GEMM (ops.fold_norm); the final norm
        + the 1x1 conv are ONE GEMM whose epilogue drops the CLS row
    """

    def __init__(self, sd: Dict[str, Tensor], cfg, device, conv_w: Tensor, conv_b: Tensor, prefix: str = ""):
        self.cfg, self.device = cfg, device
        self.fold = FOLD_NORM_IMAGE
        D, p = cfg.embed_dim, cfg.patch_size
        self.heads, self.hd = cfg.num_heads, cfg.embed_dim // cfg.num_heads
        if self.hd != 64:
            raise ValueError(f"DINOv2 attention kernel is instantiated for head_dim 64 (ViT-L), got {self.hd}")
        d = lambda k: _dev(sd[prefix + k], device)
        f32 = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)
        K = 3 * p * p
        self.k_pad = (K + 1 + 7) // 8 * 8
        w = torch.zeros(D, self.k_pad, dtype=BF16, device=device)
        w[:, :K] = d("patch_embed.proj.weight").reshape(D, K)
        self.w_patch = w          # column K (the CLS marker column of ops.patchify) stays zero
        pos = _resample_pos_embed(f32("pos_embed"), cfg.grid, cfg.interpolate_offset)
        pos[0] += f32("cls_token").reshape(D)
        pos[1:] += f32("patch_embed.proj.bias")
        self.pos = pos.to(BF16).contiguous()
        self.layers = []
        for i in range(cfg.depth):
            bp = f"blocks.{i}."
            ls1, ls2 = f32(bp + "ls1.gamma"), f32(bp + "ls2.gamma")
            scaled = lambda name, ls: ((f32(name + ".weight") * ls[:, None]).to(BF16).contiguous(),
                                       (f32(name + ".bias") * ls).to(BF16).contiguous())
            L = dict(ln1=(d(bp + "norm1.weight"), d(bp + "norm1.bias")), ln2=(d(bp + "norm2.weight"), d(bp + "norm2.bias")),
                     w_qkv=d(bp + "attn.qkv.weight"), b_qkv=d(bp + "attn.qkv.bias"),
                     w1=d(bp + "mlp.fc1.weight"), b1=d(bp + "mlp.fc1.bias"))
            L["w_o"], L["b_o"] = scaled(bp + "attn.proj", ls1)
            L["w2"], L["b2"] = scaled(bp + "mlp.fc2", ls2)
            if self.fold:
                L["f_qkv"] = ops.fold_norm(L["w_qkv"], L["ln1"][0], L["ln1"][1], L["b_qkv"])
                L["f_1"] = ops.fold_norm(L["w1"], L["ln2"][0], L["ln2"][1], L["b1"])
                del L["w_qkv"], L["w1"]
            self.layers.append(L)
        self.norm = (d("norm.weight"), d("norm.bias"))
        self.w_conv = _dev(conv_w, device).reshape(conv_w.shape[0], D).contiguous()
        self.b_conv = _dev(conv_b, device)
        if self.fold:
            self.f_conv = ops.fold_norm(self.w_conv, self.norm[0], self.norm[1], self.b_conv)
        self.scratch = _Scratch(device)
        self._drop_cls: Dict[int, Tensor] = {}

    def _drop_cls_map(self, B: int, as_src: bool = False) -> Tensor:
        key = (B, as_src)
        if key not in self._drop_cls:
            T = self.cfg.grid ** 2 + 1
            if as_src:   # output row -> source row (for the gathering norm kernel)
                r = torch.arange(B * (T - 1))
                m = (r // (T - 1)) * T + r % (T - 1) + 1
            else:        # source row -> output row, CLS rows dropped (for the GEMM epilogue)
                t = torch.arange(B * T)
                n, s = t // T, t % T
                m = torch.where(s > 0, n * (T - 1) + s - 1, torch.full_like(t, -1))
            self._drop_cls[key] = m.to(torch.int32).to(self.device)
        return self._drop_cls[key]

    def _blocks(self, images: Tensor):
        cfg = self.cfg
        B = images.shape[0]
        T, D, H, hd = cfg.grid ** 2 + 1, cfg.embed_dim, self.heads, self.hd
        if images.shape[-1] != cfg.img_size or images.shape[-2] != cfg.img_size:
            raise ValueError(f"DINOv2 encoder is laid out for {cfg.img_size}x{cfg.img_size} images, got {tuple(images.shape)}")
        T_pad = (T + 7) // 8 * 8
        a = ops.patchify(images.contiguous(), cfg.patch_size, self.k_pad, cls_rows=1)
        stA = self.scratch.gemm_stats(B * T, D, cfg.ln_eps, "a") if self.fold else None
        stB = self.scratch.gemm_stats(B * T, D, cfg.ln_eps, "b") if self.fold else None
        x = ops.gemm(a, self.w_patch, None, residual=self.pos, res_mod=T, stats_out=stA)
        del a
        q = self.scratch.zeros("q", B * H, T_pad, hd)
        k = self.scratch.zeros("k", B * H, T_pad, hd)
        vt = self.scratch.zeros("vt", B * H, hd, T_pad)
        for L in self.layers:
            if self.fold:
                wq, bq = L["f_qkv"]
                ops.gemm_qkv(x, wq, bq, q, k, vt, heads=H, head_dim=hd, seq_in=T, seq_pad=T_pad, row_stats=stA)
                h = self.scratch.zeros("o", B * T, D)
            else:
                h = ops.layernorm(x, L["ln1"][0], L["ln1"][1], cfg.ln_eps)
                ops.gemm_qkv(h, L["w_qkv"], L["b_qkv"], q, k, vt, heads=H, head_dim=hd, seq_in=T, seq_pad=T_pad)
            ops.attention(q, k, vt, h, batch=B, heads=H, head_dim=hd, seq=T, seq_pad=T_pad, scale=hd ** -0.5)
            ops.gemm(h, L["w_o"], L["b_o"], residual=x, out=x, stats_out=stB)
            if self.fold:
                w1, b1 = L["f_1"]
                m = ops.gemm(x, w1, b1, act="gelu", row_stats=stB)
            else:
                h = ops.layernorm(x, L["ln2"][0], L["ln2"][1], cfg.ln_eps)
                m = ops.gemm(h, L["w1"], L["b1"], act="gelu")
            ops.gemm(m, L["w2"], L["b2"], residual=x, out=x, stats_out=stA)
            del m
        return x, stA, B, T

    def patch_tokens(self, images: Tensor) -> Tensor:
        """`x_norm_patchtokens` [B, g*g, embed_dim] bf16 (what get_dinov2_visual_embs reshapes, LISA.py:192-195)."""
        x, _, B, T = self._blocks(images)
        y = ops.layernorm(x, self.norm[0], self.norm[1], self.cfg.ln_eps, src_row_map=self._drop_cls_map(B, True),
                          rows_out=B * (T - 1))
        return y.view(B, T - 1, -1)

    def forward(self, images: Tensor, private: Optional[dict] = None) -> Tensor:
        """[B,3,S,S] bf16 -> token-major `lisa_dino_conv` output [B, g*g, out_chans] bf16 (NHWC)."""
        x, stA, B, T = self._blocks(images)
        if self.fold:
            w, b = self.f_conv
            y = ops.gemm(x, w, b, row_stats=stA, out_row_map=self._drop_cls_map(B), out_rows=B * (T - 1))
        else:
            y = ops.layernorm(x, self.norm[0], self.norm[1], self.cfg.ln_eps)
            y = ops.gemm(y, self.w_conv, self.b_conv, out_row_map=self._drop_cls_map(B), out_rows=B * (T - 1))
        return y.view(B, T - 1, -1)


# ==============================================================================================
# CLIP ViT-L/14 tower + mm_projector
# ==============================================================================================
class ClipTower:
    def __init__(self, sd: Dict[str, Tensor], cfg, device, proj_w: Tensor, proj_b: Tensor,
                 prefix: str = "vision_model."):
        self.cfg, self.device = cfg, device
        self.fold = FOLD_NORM_TEXT
        D, p = cfg.hidden, cfg.patch_size
        self.heads, self.hd = cfg.heads, cfg.hidden // cfg.heads
        if self.hd != 64:
            raise ValueError(f"CLIP attention kernel is instantiated for head_dim 64, got {self.hd}")
        d = lambda k: _dev(sd[prefix + k], device)
        K = 3 * p * p
        self.k_pad = (K + 1 + 7) // 8 * 8
        w = torch.zeros(D, self.k_pad, dtype=BF16, device=device)
        w[:, :K] = d("embeddings.patch_embedding.weight").reshape(D, K)
        w[:, K] = d("embeddings.class_embedding")  # selected by the CLS row's one-hot column
        self.w_patch = w
        self.pos = d("embeddings.position_embedding.weight")
        self.pre_ln = (d("pre_layrnorm.weight"), d("pre_layrnorm.bias"))
        n_run = cfg.layers + 1 + cfg.select_layer if cfg.select_layer < 0 else cfg.select_layer
        self.layers = []
        for i in range(n_run):
            lp = f"encoder.layers.{i}."
            L = dict(
                ln1=(d(lp + "layer_norm1.weight"), d(lp + "layer_norm1.bias")),
                ln2=(d(lp + "layer_norm2.weight"), d(lp + "layer_norm2.bias")),
                w_qkv=torch.cat([d(lp + f"self_attn.{n}_proj.weight") for n in "qkv"], 0).contiguous(),
                b_qkv=torch.cat([d(lp + f"self_attn.{n}_proj.bias") for n in "qkv"], 0).contiguous(),
                w_o=d(lp + "self_attn.out_proj.weight"), b_o=d(lp + "self_attn.out_proj.bias"),
                w1=d(lp + "mlp.fc1.weight"), b1=d(lp + "mlp.fc1.bias"),
                w2=d(lp + "mlp.fc2.weight"), b2=d(lp + "mlp.fc2.bias"),
            )
            if self.fold:
                L["f_qkv"] = ops.fold_norm(L["w_qkv"], L["ln1"][0], L["ln1"][1], L["b_qkv"])
                L["f_1"] = ops.fold_norm(L["w1"], L["ln2"][0], L["ln2"][1], L["b1"])
                del L["w_qkv"], L["w1"]
            self.layers.append(L)
        self.proj_w, self.proj_b = _dev(proj_w, device), _dev(proj_b, device)
        self.scratch = _Scratch(device)
        self._drop_cls: Dict[int, Tensor] = {}

    def _drop_cls_map(self, N: int) -> Tensor:
        if N not in self._drop_cls:
            T = self.cfg.tokens
            t = torch.arange(N * T)
            n, s = t // T, t % T
            m = torch.where(s > 0, n * (T - 1) + s - 1, torch.full_like(t, -1))
            self._drop_cls[N] = m.to(torch.int32).to(self.device)
        return self._drop_cls[N]

    def forward(self, images_clip: Tensor) -> Tensor:
        """[N,3,224,224] bf16 -> projected patch features [N, 256, proj_dim] bf16."""
        cfg = self.cfg
        N = images_clip.shape[0]
        T, D, H, hd = cfg.tokens, cfg.hidden, self.heads, self.hd
        T_pad = (T + 7) // 8 * 8
        a = ops.patchify(images_clip.contiguous(), cfg.patch_size, self.k_pad, cls_rows=1)
        x = ops.gemm(a, self.w_patch, None, residual=self.pos, res_mod=T)
        x = ops.layernorm(x, self.pre_ln[0], self.pre_ln[1], cfg.eps)
        q = self.scratch.zeros("q", N * H, T_pad, hd)
        k = self.scratch.zeros("k", N * H, T_pad, hd)
        vt = self.scratch.zeros("vt", N * H, hd, T_pad)
        stB = self.scratch.gemm_stats(N * T, D, cfg.eps, "b") if self.fold else None
        st = None
        for L in self.layers:
            if self.fold:
                if st is None:  # first layer: x comes out of pre_layrnorm, not out of a GEMM
                    st = ops.norm_stats(x, cfg.eps, out=self.scratch.stats(N * T, 1, "n"))
                wq, bq = L["f_qkv"]
                ops.gemm_qkv(x, wq, bq, q, k, vt, heads=H, head_dim=hd, seq_in=T, seq_pad=T_pad, row_stats=st)
                h = self.scratch.zeros("o", N * T, D)
            else:
                h = ops.layernorm(x, L["ln1"][0], L["ln1"][1], cfg.eps)
                ops.gemm_qkv(h, L["w_qkv"], L["b_qkv"], q, k, vt, heads=H, head_dim=hd, seq_in=T, seq_pad=T_pad)
            ops.attention(q, k, vt, h, batch=N, heads=H, head_dim=hd, seq=T, seq_pad=T_pad, scale=hd ** -0.5)
            ops.gemm(h, L["w_o"], L["b_o"], residual=x, out=x, stats_out=stB)
            if self.fold:
                w1, b1 = L["f_1"]
                m = ops.gemm(x, w1, b1, act="quick_gelu", row_stats=stB)
                st = self.scratch.gemm_stats(N * T, D, cfg.eps, "a")
            else:
                h = ops.layernorm(x, L["ln2"][0], L["ln2"][1], cfg.eps)
                m = ops.gemm(h, L["w1"], L["b1"], act="quick_gelu")
            ops.gemm(m, L["w2"], L["b2"], residual=x, out=x, stats_out=st)
        feats = ops.gemm(x, self.proj_w, self.proj_b, out_row_map=self._drop_cls_map(N), out_rows=N * (T - 1))
        return feats.view(N, T - 1, -1)


# ==============================================================================================
# LLaMA decoder stack
# ==============================================================================================
class LlamaDecoder:
    def __init__(self, sd: Dict[str, Tensor], cfg, device, prefix: str = "", max_seq: int = 1024,
                 lm_head: Optional[Tensor] = None):
        self.cfg, self.device = cfg, device
        self.fold = FOLD_NORM_TEXT
        if cfg.head_dim != 128:
            raise ValueError(f"LLaMA attention kernel is instantiated for head_dim 128, got {cfg.head_dim}")
        d = lambda k: _dev(sd[prefix + k], device)
        self.embed = d("embed_tokens.weight")
        self.norm = d("norm.weight")
        self.layers = []
        for i in range(cfg.layers):
            lp = f"layers.{i}."
            gate, up = d(lp + "mlp.gate_proj.weight"), d(lp + "mlp.up_proj.weight")
            L = dict(
                rms1=d(lp + "input_layernorm.weight"), rms2=d(lp + "post_attention_layernorm.weight"),
                w_qkv=torch.cat([d(lp + f"self_attn.{n}_proj.weight") for n in "qkv"], 0).contiguous(),
                w_o=d(lp + "self_attn.o_proj.weight"),
                # rows interleaved (gate0, up0, gate1, up1, ...) for the SwiGLU epilogue
                w_gu=torch.stack([gate, up], dim=1).reshape(2 * cfg.mlp, cfg.hidden).contiguous(),
                w_down=d(lp + "mlp.down_proj.weight"),
            )
            del gate, up
            if self.fold:  # RMSNorm: gamma folds into the weight, rstd is applied per row in the epilogue
                L["w_qkv"] = ops.fold_norm(L["w_qkv"], L["rms1"], rms=True)[0]
                L["w_gu"] = ops.fold_norm(L["w_gu"], L["rms2"], rms=True)[0]
            self.layers.append(L)
        # lm_head is dead work at inference (LISA.py:283,318); the training forward needs it.  Rows padded to a
        # multiple of 8 (vocab 32003 -> 32008, zero rows; the CE kernel reads the first `vocab` columns only).
        self.w_lm = None
        if lm_head is not None:
            v_pad = (lm_head.shape[0] + 7) // 8 * 8
            w = torch.zeros(v_pad, cfg.hidden, dtype=BF16, device=device)
            w[:lm_head.shape[0]] = _dev(lm_head, device)
            self.w_lm = ops.fold_norm(w, self.norm, rms=True)[0] if self.fold else w
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.float32) / cfg.head_dim))
        fr = torch.outer(torch.arange(max_seq, dtype=torch.float32), inv)
        self.rope_cos = fr.cos().to(BF16).to(device).contiguous()
        self.rope_sin = fr.sin().to(BF16).to(device).contiguous()
        self.max_seq = max_seq
        self.scratch = _Scratch(device)

    def forward(self, embeds: Tensor, n_seq: int, T: int, kv_len: Optional[Tensor],
                out_rows: Optional[Tensor] = None, with_logits: bool = False):
        """embeds [n_seq*T, hidden] bf16 -> final-norm hidden states; with `out_rows` (int32 flat row
        indices) only those rows are normalised and returned (the row-wise norm commutes with the gather).
        with_logits (training forward, llava_llama.py:104-105): also returns lm_head(norm(x)) for every row,
        bf16 [n_seq*T, vocab padded to 8] — the final norm folded into the lm_head GEMM."""
        cfg = self.cfg
        if with_logits and self.w_lm is None:
            raise RuntimeError("the training forward needs `lm_head.weight` in the state dict")
        if T > self.max_seq:
            raise ValueError(f"sequence length {T} exceeds the RoPE table ({self.max_seq})")
        H, hd = cfg.heads, cfg.head_dim
        T_pad = (T + 7) // 8 * 8
        q = self.scratch.zeros("q", n_seq * H, T_pad, hd)
        k = self.scratch.zeros("k", n_seq * H, T_pad, hd)
        vt = self.scratch.zeros("vt", n_seq * H, hd, T_pad)
        x = embeds
        scale = 1.0 / math.sqrt(hd)
        stB = self.scratch.gemm_stats(n_seq * T, cfg.hidden, cfg.eps, "b", rms=True) if self.fold else None
        st = None
        for L in self.layers:
            if self.fold:
                if st is None:  # first layer: x is the spliced embedding sequence, not a GEMM output
                    st = ops.norm_stats(x, cfg.eps, rms=True, out=self.scratch.stats(n_seq * T, 1, "n"))
                ops.gemm_qkv(x, L["w_qkv"], None, q, k, vt, heads=H, head_dim=hd, seq_in=T, seq_pad=T_pad,
                             rope_cos=self.rope_cos, rope_sin=self.rope_sin, row_stats=st)
                h = self.scratch.zeros("o", n_seq * T, cfg.hidden)
            else:
                h = ops.rmsnorm(x, L["rms1"], cfg.eps)
                ops.gemm_qkv(h, L["w_qkv"], None, q, k, vt, heads=H, head_dim=hd, seq_in=T, seq_pad=T_pad,
                             rope_cos=self.rope_cos, rope_sin=self.rope_sin)
            ops.attention(q, k, vt, h, batch=n_seq, heads=H, head_dim=hd, seq=T, seq_pad=T_pad, scale=scale,
                          causal=True, kv_len=kv_len)
            if L is self.layers[-1] and out_rows is not None and not with_logits and LAST_LAYER_ROWS:
                # Only the hidden state that predicts [SEG] leaves the decoder (LISA.py:322-337), so everything
                # after the last layer's attention — o_proj, both norms, the MLP — is row-wise and runs on the
                # gathered rows alone.  Keys/values of the last layer still cover every position.
                xr = ops.gather_rows(x, out_rows)
                xr = ops.gemm(ops.gather_rows(h, out_rows), L["w_o"], None, residual=xr)
                if self.fold:
                    m = ops.gemm(xr, L["w_gu"], None, swiglu=True, row_stats=ops.norm_stats(xr, cfg.eps, rms=True))
                else:
                    m = ops.gemm(ops.rmsnorm(xr, L["rms2"], cfg.eps), L["w_gu"], None, swiglu=True)
                xr = ops.gemm(m, L["w_down"], None, residual=xr)
                return ops.rmsnorm(xr, self.norm, cfg.eps)
            ops.gemm(h, L["w_o"], None, residual=x, out=x, stats_out=stB)
            if self.fold:
                m = ops.gemm(x, L["w_gu"], None, swiglu=True, row_stats=stB)
                st = self.scratch.gemm_stats(n_seq * T, cfg.hidden, cfg.eps, "a", rms=True)
            else:
                h = ops.rmsnorm(x, L["rms2"], cfg.eps)
                m = ops.gemm(h, L["w_gu"], None, swiglu=True)
            ops.gemm(m, L["w_down"], None, residual=x, out=x, stats_out=st)
        if out_rows is not None:
            hidden = ops.rmsnorm(x, self.norm, cfg.eps, src_row_map=out_rows, rows_out=out_rows.numel())
        else:
            hidden = ops.rmsnorm(x, self.norm, cfg.eps)
        if not with_logits:
            return hidden
        if self.fold and st is not None:
            logits = ops.gemm(x, self.w_lm, None, row_stats=st)
        else:
            logits = ops.gemm(hidden if out_rows is None else ops.rmsnorm(x, self.norm, cfg.eps), self.w_lm, None)
        return hidden, logits
