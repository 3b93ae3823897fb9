"""Thin torch-tensor front-ends over the C ABI (device pointers + the current CUDA stream).

PyTorch is only the allocator / stream provider here; every function enqueues hand-written
sm_100a kernels from libllmseg_b200.so and raises if that is impossible.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import AttnParams, GemmParams, check

ACT = {None: 0, "none": 0, "gelu": 1, "quick_gelu": 2, "relu": 3}
GEMM_PLAIN, GEMM_SWIGLU, GEMM_QKV = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_bf16(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError("llmseg_b200 ops need CUDA tensors (no CPU fallback exists)")
            if t.dtype != torch.bfloat16:
                raise TypeError(f"expected bfloat16, got {t.dtype}")
            if t.stride(-1) != 1:
                raise ValueError("innermost dimension must be contiguous")


_gemm_ws = {}
USE_GEMM_WORKSPACE = True  # tests flip this to compare against the plain (no stream-K) schedule


def _workspace(p: GemmParams, device) -> None:
    """Stream-K scratch for llmseg_gemm: one zero-filled buffer per (device, stream) — GEMMs on one stream
    never overlap, so they can share it (see include/llmseg_b200.h)."""
    if not USE_GEMM_WORKSPACE:
        return
    key = (torch.device(device).index or 0, _stream())
    ws = _gemm_ws.get(key)
    if ws is None:
        ws = torch.zeros(_lib.lib().llmseg_gemm_workspace_bytes(), dtype=torch.uint8, device=device)
        _gemm_ws[key] = ws
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()


def ensure_workspace(device) -> None:
    """Allocate (outside any graph capture) the GEMM workspace of the current stream on `device`."""
    _workspace(GemmParams(), device)


class RowStats:
    """Per-row normalisation statistics for gemm(row_stats=...): either (mean, rstd) pairs from norm_stats
    (parts == 0, t = fp32 [rows, 2]) or the (sum, sum of squares) partials a previous gemm(stats_out=...)
    wrote from its epilogue (t = fp32 [rows, parts, 2]) together with what is needed to finish them."""
    __slots__ = ("t", "parts", "dim", "eps", "rms", "final")

    def __init__(self, t: torch.Tensor, parts: int, dim: int, eps: float, rms: bool,
                 final: Optional[torch.Tensor] = None):
        self.t, self.parts, self.dim, self.eps, self.rms = t, parts, dim, float(eps), bool(rms)
        self.final = final   # fp32 [rows, 2] (mean, rstd) finished in-kernel by the producing GEMM, or None


def gemm_stats_buffer(M: int, N: int, rows_out: int, eps: float, *, rms: bool = False, device="cuda",
                      out: Optional[torch.Tensor] = None) -> RowStats:
    """Buffer for gemm(..., stats_out=...) of an [M, K] x [N, K] PLAIN problem writing `rows_out` rows of
    width N: the norm that follows (over those N columns) reads its statistics from here."""
    parts = _lib.lib().llmseg_gemm_stats_parts(M, N)
    if out is None:
        out = torch.empty((rows_out, parts + 1, 2), dtype=torch.float32, device=device)
    assert out.dtype == torch.float32 and out.numel() == rows_out * (parts + 1) * 2
    flat = out.view(-1)
    part_t = flat[:rows_out * parts * 2].view(rows_out, parts, 2)
    final = flat[rows_out * parts * 2:].view(rows_out, 2) if rows_out == M else None
    return RowStats(part_t, parts, N, eps, rms, final)


def gemm_stats_parts(M: int, N: int) -> int:
    return _lib.lib().llmseg_gemm_stats_parts(M, N)


def _norm_fold_args(p: GemmParams, row_stats: Optional[RowStats], M: int) -> None:
    if row_stats is None:
        return
    t, parts = row_stats.t, row_stats.parts
    if row_stats.final is not None:   # finished by the producing GEMM: read (mean, rstd) directly
        t, parts = row_stats.final, 0
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == 2 * M * max(parts, 1)):
        raise ValueError("row_stats must hold fp32 [M, 2] or [M, parts, 2] on the GPU")
    p.row_stats = t.data_ptr()
    p.row_stats_parts, p.norm_dim, p.norm_eps, p.norm_rms = parts, row_stats.dim, row_stats.eps, int(row_stats.rms)


def fold_norm(w: torch.Tensor, gamma: torch.Tensor, beta: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, rms: bool = False):
    """Weight-side half of folding y = Norm(x) @ w.T + bias into the GEMM on the un-normalised x
    (include/llmseg_b200.h, row_stats): returns (w2 bf16 [N,K], bias2 bf16 [N] or None) with
    w2 = w * gamma, rows centred for LayerNorm (so the mean term vanishes), bias2 = bias + w @ beta.
    One-off, at model load."""
    wf = w.float() * gamma.float()[None, :]
    if not rms:
        wf = wf - wf.mean(dim=1, keepdim=True)
    bias2 = None
    if bias is not None or beta is not None:
        b = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
        if bias is not None:
            b += bias.float()
        if beta is not None:
            b += w.float() @ beta.float()
        bias2 = b.to(torch.bfloat16).contiguous()
    return wf.to(torch.bfloat16).contiguous(), bias2


def norm_stats(x: torch.Tensor, eps: float, *, rms: bool = False, out: Optional[torch.Tensor] = None) -> RowStats:
    """(mean, rstd) per row of x (rms: (0, rsqrt(mean(x^2)+eps))) as a RowStats for gemm(row_stats=...)."""
    _req_bf16(x)
    dim = x.shape[-1]
    x2 = x.reshape(-1, dim) if x.dim() != 2 else x
    if out is None:
        out = torch.empty((x2.shape[0], 2), dtype=torch.float32, device=x.device)
    check(_lib.lib().llmseg_norm_stats(x2.data_ptr(), x2.stride(0), x2.shape[0], dim, float(eps), int(rms),
                                       out.data_ptr(), _stream()), "norm_stats")
    return RowStats(out, 0, dim, eps, rms)


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
         act: Optional[str] = None, residual: Optional[torch.Tensor] = None, res_mod: int = 0,
         out: Optional[torch.Tensor] = None, out_row_map: Optional[torch.Tensor] = None,
         out_rows: Optional[int] = None, swiglu: bool = False,
         row_stats: Optional[RowStats] = None, stats_out: Optional[RowStats] = None) -> torch.Tensor:
    """out = act(a @ w.T + bias) (+ residual).  a: [M,K] bf16, w: [N,K] bf16 (nn.Linear layout).

    out_row_map (int32 [M]) scatters GEMM row r to output row out_row_map[r] (negative = dropped);
    the residual is read at the same output row (modulo res_mod when given).
    swiglu: w rows are (gate0, up0, gate1, up1, ...) and out has N/2 columns.
    row_stats: the row normalisation of `a` folded into the epilogue (w, bias from fold_norm).
    stats_out: gemm_stats_buffer(M, N, ...) to fill with this GEMM's per-row output statistics.
    """
    _req_bf16(a, w, bias, residual, out)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    n_out = N // 2 if swiglu else N
    if out is None:
        rows = out_rows if out_rows is not None else M
        out = torch.empty((rows, n_out), dtype=torch.bfloat16, device=a.device)
    p = GemmParams()
    p.M, p.N, p.K = M, N, K
    p.A, p.lda = a.data_ptr(), a.stride(0)
    p.W, p.ldw = w.data_ptr(), w.stride(0)
    p.C, p.ldc = out.data_ptr(), out.stride(0)
    p.bias = _ptr(bias)
    p.residual, p.ldr, p.res_mod = _ptr(residual), (residual.stride(0) if residual is not None else 0), res_mod
    p.act = ACT[act]
    p.mode = GEMM_SWIGLU if swiglu else GEMM_PLAIN
    p.out_row_map = _ptr(out_row_map)
    _norm_fold_args(p, row_stats, M)
    if stats_out is not None:
        if swiglu or stats_out.parts != _lib.lib().llmseg_gemm_stats_parts(M, N) or stats_out.dim != N:
            raise ValueError("stats_out must come from gemm_stats_buffer(M, N, ...) of this (plain) problem")
        p.stats_out = stats_out.t.data_ptr()
        if stats_out.final is not None:
            if out_row_map is not None or not USE_GEMM_WORKSPACE:
                raise ValueError("in-kernel statistics need the workspace and unscattered output rows")
            p.stats_final = stats_out.final.data_ptr()
            p.stats_dim, p.stats_eps, p.stats_rms = stats_out.dim, stats_out.eps, int(stats_out.rms)
    _workspace(p, a.device)
    check(_lib.lib().llmseg_gemm(C.byref(p), _stream()), "gemm")
    return out


def gemm_qkv(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], q: torch.Tensor,
             k: torch.Tensor, vt: torch.Tensor, *, heads: int, head_dim: int, seq_in: int,
             seq_pad: int, rope_cos: Optional[torch.Tensor] = None,
             rope_sin: Optional[torch.Tensor] = None, row_map: Optional[torch.Tensor] = None,
             row_stats: Optional[RowStats] = None) -> None:
    """QKV projection writing q,k [(b*heads+h), seq_pad, hd] and vt [(b*heads+h), hd, seq_pad].
    row_map (int32 [M]): GEMM row r lands at position m = row_map[r] -> (b, s) = divmod(m, seq_in)."""
    _req_bf16(a, w, bias, q, k, vt, rope_cos, rope_sin)
    M, K = a.shape
    p = GemmParams()
    p.M, p.N, p.K = M, w.shape[0], K
    p.A, p.lda = a.data_ptr(), a.stride(0)
    p.W, p.ldw = w.data_ptr(), w.stride(0)
    p.bias = _ptr(bias)
    p.mode = GEMM_QKV
    p.q, p.k, p.vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
    p.heads, p.head_dim, p.seq_in, p.seq_pad = heads, head_dim, seq_in, seq_pad
    p.rope_cos, p.rope_sin = _ptr(rope_cos), _ptr(rope_sin)
    p.out_row_map = _ptr(row_map)
    _norm_fold_args(p, row_stats, M)
    _workspace(p, a.device)
    check(_lib.lib().llmseg_gemm(C.byref(p), _stream()), "gemm_qkv")


def fill_kv_rows(k: torch.Tensor, vt: torch.Tensor, bias_qkv: torch.Tensor, pos_map: torch.Tensor, *, batch: int,
                 heads: int, head_dim: int, seq_in: int, seq_pad: int, seq_ids: Optional[torch.Tensor] = None) -> None:
    """k / vt entries of window-padding positions (pos_map < 0) <- projection bias (zero tokens after LN).
    seq_ids (int32): the sequences that contain padding; None visits all `batch` sequences."""
    _req_bf16(k, vt, bias_qkv)
    assert pos_map.dtype == torch.int32 and pos_map.numel() == batch * seq_in
    n = batch
    if seq_ids is not None:
        assert seq_ids.dtype == torch.int32 and seq_ids.is_cuda
        n = seq_ids.numel()
        if n == 0:
            return
    check(_lib.lib().llmseg_fill_kv_rows(k.data_ptr(), vt.data_ptr(), bias_qkv.data_ptr(), pos_map.data_ptr(),
                                         _ptr(seq_ids), n, heads, head_dim, seq_in, seq_pad, _stream()),
          "fill_kv_rows")


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, *, batch: int,
              heads: int, head_dim: int, seq: int, seq_pad: int, scale: float, causal: bool = False,
              kv_len: Optional[torch.Tensor] = None, qext: Optional[torch.Tensor] = None,
              kext: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
              ext_cols: int = 0, out_row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused attention over the q/k/vt buffers of gemm_qkv; out is [batch*seq, heads*head_dim]."""
    _req_bf16(q, k, vt, out, qext, kext, row_bias)
    p = AttnParams()
    p.q, p.k, p.vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
    p.out, p.ldo = out.data_ptr(), out.stride(0)
    p.batch, p.heads, p.head_dim, p.seq, p.seq_pad = batch, heads, head_dim, seq, seq_pad
    p.scale, p.causal = float(scale), int(causal)
    p.kv_len = _ptr(kv_len)
    p.ext_cols, p.qext, p.kext, p.row_bias = ext_cols, _ptr(qext), _ptr(kext), _ptr(row_bias)
    p.out_row_map = _ptr(out_row_map)
    check(_lib.lib().llmseg_attention(C.byref(p), _stream()), "attention")
    return out


def relpos_prep(q: torch.Tensor, rel_hw: torch.Tensor, *, bh: int, seq: int, seq_pad: int,
                head_dim: int, grid: int, inv_scale: float, qext: torch.Tensor,
                row_bias: Optional[torch.Tensor] = None) -> None:
    """qext/row_bias <- gathered q·rel_posᵀ (decomposed rel-pos, see include/llmseg_b200.h)."""
    _req_bf16(q, rel_hw, qext, row_bias)
    check(_lib.lib().llmseg_relpos_prep(q.data_ptr(), rel_hw.data_ptr(), rel_hw.shape[0], bh, seq,
                                        seq_pad, head_dim, grid, float(inv_scale), qext.data_ptr(),
                                        qext.shape[-1], _ptr(row_bias), _stream()), "relpos_prep")


def make_kext(grid: int, device) -> torch.Tensor:
    """Constant one-hot key-position matrix for the rel-pos score extension."""
    if grid == 14:
        e = torch.zeros(256, 32)
        key = torch.arange(196)
        e[key, key // 14] = 1
        e[key, 14 + key % 14] = 1
    elif grid == 64:
        e = torch.zeros(128, 64)
        key = torch.arange(128)
        e[key, key % 64] = 1
    else:
        raise ValueError(f"rel-pos extension supports grid 14 (windows) or 64 (global), got {grid}")
    return e.to(device=device, dtype=torch.bfloat16)


def make_rel_hw(rel_h: torch.Tensor, rel_w: torch.Tensor) -> torch.Tensor:
    """Stack rel_pos_h / rel_pos_w ([2g-1, hd] each) into the zero-padded table relpos_prep reads."""
    t = rel_h.shape[0]
    if t == 27:      # 14x14 windows: one 128-column GEMM tile, rel_w at row 32
        n_pad, w0 = 64, 32
    elif t == 127:   # 64x64 global: rel_h -> N-tile 0, rel_w -> N-tile 1
        n_pad, w0 = 256, 128
    else:
        raise ValueError(f"rel-pos tables must have 27 (window 14) or 127 (grid 64) rows, got {t}")
    out = torch.zeros(n_pad, rel_h.shape[1], dtype=torch.bfloat16, device=rel_h.device)
    out[:t] = rel_h
    out[w0:w0 + t] = rel_w
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *,
              src_row_map: Optional[torch.Tensor] = None, rows_out: Optional[int] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_bf16(x, gamma, beta, out)
    dim = x.shape[-1]
    x2 = x.reshape(-1, dim) if x.dim() != 2 else x
    rows = rows_out if rows_out is not None else x2.shape[0]
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().llmseg_layernorm(x2.data_ptr(), x2.stride(0), out.data_ptr(), out.stride(0),
                                      gamma.data_ptr(), beta.data_ptr(), rows, dim, float(eps),
                                      _ptr(src_row_map), _stream()), "layernorm")
    return out


def rmsnorm(x: torch.Tensor, gamma: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None, *,
            src_row_map: Optional[torch.Tensor] = None, rows_out: Optional[int] = None) -> torch.Tensor:
    _req_bf16(x, gamma, out)
    dim = x.shape[-1]
    x2 = x.reshape(-1, dim) if x.dim() != 2 else x
    rows = rows_out if rows_out is not None else x2.shape[0]
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().llmseg_rmsnorm(x2.data_ptr(), x2.stride(0), out.data_ptr(), out.stride(0),
                                    gamma.data_ptr(), rows, dim, float(eps), _ptr(src_row_map),
                                    _stream()), "rmsnorm")
    return out


def patchify(images: torch.Tensor, patch: int, k_pad: int, cls_rows: int = 0) -> torch.Tensor:
    """[B,3,S,S] bf16 NCHW -> [B*(g*g+cls_rows), k_pad] patch rows (conv-as-GEMM operand)."""
    _req_bf16(images)
    B, Cc, S, S2 = images.shape
    assert Cc == 3 and S == S2 and images.is_contiguous()
    g = S // patch
    out = torch.empty((B * (g * g + cls_rows), k_pad), dtype=torch.bfloat16, device=images.device)
    check(_lib.lib().llmseg_patchify(images.data_ptr(), out.data_ptr(), B, S, patch, k_pad, cls_rows,
                                     _stream()), "patchify")
    return out


def embed_splice(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], embed: torch.Tensor,
                 feats: torch.Tensor, *, image_token: int, seg_token: int,
                 seg_row_out: Optional[torch.Tensor] = None):
    """-> (embeds [N*T, D] bf16, kv_len int32 [N], seg_row int32 [N]) with T = T_text + F - 1."""
    _req_bf16(embed, feats)
    assert input_ids.dtype == torch.int64 and input_ids.is_cuda and input_ids.is_contiguous()
    N, Tt = input_ids.shape
    F_, D = feats.shape[-2], feats.shape[-1]
    T = Tt + F_ - 1
    m = None
    if attention_mask is not None:
        m = attention_mask.to(torch.uint8).contiguous()
    out = torch.empty((N * T, D), dtype=torch.bfloat16, device=embed.device)
    kv_len = torch.empty(N, dtype=torch.int32, device=embed.device)
    seg_row = seg_row_out if seg_row_out is not None else torch.empty(N, dtype=torch.int32, device=embed.device)
    assert seg_row.dtype == torch.int32 and seg_row.numel() == N and seg_row.is_cuda
    check(_lib.lib().llmseg_embed_splice(input_ids.data_ptr(), _ptr(m), embed.data_ptr(), feats.data_ptr(),
                                         out.data_ptr(), kv_len.data_ptr(), seg_row.data_ptr(), N, Tt, F_,
                                         D, image_token, seg_token, embed.shape[0], _stream()),
          "embed_splice")
    return out, kv_len, seg_row


def add_rows_bcast(x: torch.Tensor, y: torch.Tensor, *, group: int = 0,
                   row_group: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_bf16(x, y)
    out = torch.empty_like(x)
    check(_lib.lib().llmseg_add_rows_bcast(x.data_ptr(), y.data_ptr(), out.data_ptr(), x.shape[0],
                                           x.shape[1], group, _ptr(row_group), _stream()),
          "add_rows_bcast")
    return out


def gather_rows(x: torch.Tensor, src_row_map: torch.Tensor) -> torch.Tensor:
    """out[r] = x[src_row_map[r]] (int32 map on the device; a negative index gives a zero row)."""
    _req_bf16(x)
    assert x.dim() == 2 and src_row_map.dtype == torch.int32 and src_row_map.is_cuda
    out = torch.empty((src_row_map.numel(), x.shape[1]), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().llmseg_gather_rows(x.data_ptr(), x.stride(0), out.data_ptr(), out.shape[0], x.shape[1],
                                        src_row_map.data_ptr(), _stream()), "gather_rows")
    return out


def im2col3x3(x: torch.Tensor, batch: int, height: int, width: int) -> torch.Tensor:
    """token-major NHWC [B*H*W, C] -> [B*H*W, 9*C] (zero padded 3x3 neighbourhoods, (ky,kx,c) order)."""
    _req_bf16(x)
    Cc = x.shape[1]
    out = torch.empty((x.shape[0], 9 * Cc), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().llmseg_im2col3x3(x.data_ptr(), out.data_ptr(), batch, height, width, Cc, _stream()),
          "im2col3x3")
    return out


def maskpool(segs: torch.Tensor, emb_tokens: torch.Tensor, mask_image: torch.Tensor) -> torch.Tensor:
    """segs [n_masks,256,256] bf16, emb_tokens [B,4096,256] bf16, mask_image int32 [n_masks] -> [n_masks,256]."""
    _req_bf16(segs, emb_tokens)
    n = segs.shape[0]
    assert segs.shape[1:] == (256, 256) and segs.is_contiguous() and emb_tokens.is_contiguous()
    ws = torch.empty(int(_lib.lib().llmseg_maskpool_workspace(n)), dtype=torch.uint8, device=segs.device)
    out = torch.empty((n, 256), dtype=torch.bfloat16, device=segs.device)
    check(_lib.lib().llmseg_maskpool(segs.data_ptr(), emb_tokens.data_ptr(), mask_image.data_ptr(), n,
                                     out.data_ptr(), ws.data_ptr(), _stream()), "maskpool")
    return out


def small_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_off: torch.Tensor,
                    kv_off: torch.Tensor, *, batch: int, heads: int, max_kv: int) -> torch.Tensor:
    """q/k/v are (possibly strided column views of) [rows, heads*32] bf16; offsets int32 [batch+1]."""
    _req_bf16(q, k, v)
    out = torch.empty((q.shape[0], heads * 32), dtype=torch.bfloat16, device=q.device)
    check(_lib.lib().llmseg_small_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
                                            v.data_ptr(), v.stride(0), out.data_ptr(), out.stride(0),
                                            q_off.data_ptr(), kv_off.data_ptr(), batch, heads, 32, max_kv,
                                            _stream()), "small_attention")
    return out


def select(feat: torch.Tensor, text: torch.Tensor, h_iou: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
           k_off: torch.Tensor, *, batch: int, k_stride: int, conv_group: Optional[torch.Tensor] = None,
           conv_valid: Optional[torch.Tensor] = None):
    """text [n_conv,256]: one row per conversation; conv_group int32 [n_conv] maps a conversation to its group
    of mask tokens (None: n_conv == batch, identity); conv_valid int32 [n_conv] (< 0: no [SEG], NaN sentinel).
    -> (sim fp32 [n_conv,k_stride], iou fp32 [batch,k_stride], best int32 [batch])."""
    _req_bf16(feat, text, h_iou, w2, b2)
    dev = feat.device
    n_conv = text.shape[0]
    for t in (conv_group, conv_valid):
        if t is not None and not (t.is_cuda and t.dtype == torch.int32 and t.numel() == n_conv):
            raise ValueError("conv_group / conv_valid must be int32 [n_conv] on the GPU")
    sim = torch.empty((n_conv, k_stride), dtype=torch.float32, device=dev)
    iou = torch.empty((batch, k_stride), dtype=torch.float32, device=dev)
    best = torch.empty(batch, dtype=torch.int32, device=dev)
    check(_lib.lib().llmseg_select(feat.data_ptr(), text.data_ptr(), h_iou.data_ptr(), w2.data_ptr(),
                                   b2.data_ptr(), k_off.data_ptr(), batch, k_stride, _ptr(conv_group),
                                   _ptr(conv_valid), n_conv, sim.data_ptr(), iou.data_ptr(), best.data_ptr(),
                                   _stream()), "select")
    return sim, iou, best


def align_iou_loss(sim: torch.Tensor, pred_iou: torch.Tensor, gt_iou: torch.Tensor,
                   temperature: float = 0.05) -> torch.Tensor:
    """-> fp32 [2] = {softmax_align_loss, iou_regression_loss} (reference model/loss.py:50-94)."""
    for t in (sim, pred_iou, gt_iou):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    out = torch.empty(2, dtype=torch.float32, device=sim.device)
    check(_lib.lib().llmseg_align_iou_loss(sim.data_ptr(), pred_iou.data_ptr(), gt_iou.data_ptr(),
                                           sim.numel(), float(temperature), out.data_ptr(), _stream()),
          "align_iou_loss")
    return out


def selector_losses(sim: torch.Tensor, pred_iou: torch.Tensor, gt_iou: torch.Tensor, gt_iop: torch.Tensor,
                    k_off: torch.Tensor, group_weight: torch.Tensor, *, ce: Optional[torch.Tensor] = None,
                    weights=(1.0, 1.0, 1.0), temperature: float = 0.05):
    """Training-forward loss assembly (reference LISA.py:416-474, loss.py:50-94) over G (image, round) groups:
    sim/pred_iou/gt_iou/gt_iop fp32 [G, k_stride], k_off int32 [G+1], group_weight fp32 [G], ce fp32 [>=1] or None,
    weights = (ce, align, regression).  -> (out4 = {loss, ce, align, regression}, per_group fp32 [G,2])."""
    G, ks = sim.shape
    for t in (sim, pred_iou, gt_iou, gt_iop):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == (G, ks)
    assert k_off.dtype == torch.int32 and k_off.numel() == G + 1 and group_weight.dtype == torch.float32
    per_group = torch.empty((G, 2), dtype=torch.float32, device=sim.device)
    out = torch.empty(4, dtype=torch.float32, device=sim.device)
    check(_lib.lib().llmseg_selector_losses(sim.data_ptr(), pred_iou.data_ptr(), gt_iou.data_ptr(), gt_iop.data_ptr(),
                                            k_off.data_ptr(), G, ks, float(temperature), group_weight.data_ptr(),
                                            _ptr(ce), float(weights[0]), float(weights[1]), float(weights[2]),
                                            per_group.data_ptr(), out.data_ptr(), _stream()), "selector_losses")
    return out, per_group


def lm_cross_entropy(logits: torch.Tensor, input_ids: torch.Tensor, labels: torch.Tensor, *, n_img_tokens: int,
                     vocab: int, image_token: int, ignore_index: int = -100):
    """Shifted LM cross entropy on spliced labels (reference llava_llama.py:107-118, llava_arch.py:185-245).
    logits bf16 [N*T, >=vocab] with T = T_text + n_img_tokens - 1.  -> (out2 = {mean CE, #targets}, row_loss [N*T])."""
    _req_bf16(logits)
    N, Tt = input_ids.shape
    T = Tt + n_img_tokens - 1
    assert logits.shape[0] == N * T and labels.shape == input_ids.shape
    assert input_ids.dtype == torch.int64 and labels.dtype == torch.int64 and input_ids.is_cuda and labels.is_cuda
    input_ids, labels = input_ids.contiguous(), labels.contiguous()
    row_loss = torch.empty(N * T, dtype=torch.float32, device=logits.device)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    check(_lib.lib().llmseg_lm_cross_entropy(logits.data_ptr(), logits.stride(0), input_ids.data_ptr(), labels.data_ptr(),
                                             N, Tt, n_img_tokens, vocab, image_token, ignore_index, row_loss.data_ptr(),
                                             out.data_ptr(), _stream()), "lm_cross_entropy")
    return out, row_loss


def dice_bce_loss(logits: torch.Tensor, targets: torch.Tensor, num_masks: float) -> torch.Tensor:
    """logits/targets fp32 [n,H,W] -> fp32 [2] = {dice_loss, sigmoid_ce_loss} (reference model/loss.py:4-47)."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    assert targets.shape == logits.shape and targets.dtype == torch.float32 and targets.is_contiguous()
    n = logits.shape[0]
    hw = logits[0].numel()
    ws = torch.empty(2 * n, dtype=torch.float32, device=logits.device)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    check(_lib.lib().llmseg_dice_bce_loss(logits.data_ptr(), targets.data_ptr(), n, hw, float(num_masks),
                                          ws.data_ptr(), out.data_ptr(), _stream()), "dice_bce_loss")
    return out


# ---- SAM-Everything proposal generation (csrc/amg.cu; include/llmseg_b200.h) ---------------------------------
def _req_f32(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError("expected a contiguous fp32 CUDA tensor")


def _req_i32(t):
    if t is not None and not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise TypeError("expected a contiguous int32 CUDA tensor")


def point_tokens(points: torch.Tensor, gauss: torch.Tensor, out_tokens: torch.Tensor, point_embed: torch.Tensor,
                 not_a_point: torch.Tensor, img_size: float) -> torch.Tensor:
    """points fp32 [P,2] (x, y) -> decoder tokens bf16 [P*7, 256] (iou, 4 mask tokens, point, padding point)."""
    _req_f32(points, gauss)
    _req_bf16(out_tokens, point_embed, not_a_point)
    P = points.shape[0]
    out = torch.empty((P * 7, 256), dtype=torch.bfloat16, device=points.device)
    check(_lib.lib().llmseg_point_tokens(points.data_ptr(), P, gauss.data_ptr(), out_tokens.data_ptr(),
                                         point_embed.data_ptr(), not_a_point.data_ptr(), float(img_size), out.data_ptr(),
                                         _stream()), "point_tokens")
    return out


def tok2img_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_prompts: int, shared_kv: bool) -> torch.Tensor:
    """q [P*7, 128] (view), k / v [4096 or P*4096, 128] (column views of a wider buffer allowed) -> [P*7, 128]."""
    _req_bf16(q, k, v)
    out = torch.empty((n_prompts * 7, 128), dtype=torch.bfloat16, device=q.device)
    kbs = 0 if shared_kv else 4096 * k.stride(0)
    vbs = 0 if shared_kv else 4096 * v.stride(0)
    check(_lib.lib().llmseg_tok2img_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), kbs, v.data_ptr(),
                                              v.stride(0), vbs, out.data_ptr(), out.stride(0), n_prompts, _stream()),
          "tok2img_attention")
    return out


def img2tok_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_prompts: int, shared_q: bool) -> torch.Tensor:
    """q [4096 or P*4096, 128] (view), k / v [P*7, 128] -> [P*4096, 128]."""
    _req_bf16(q, k, v)
    out = torch.empty((n_prompts * 4096, 128), dtype=torch.bfloat16, device=q.device)
    qbs = 0 if shared_q else 4096 * q.stride(0)
    check(_lib.lib().llmseg_img2tok_attention(q.data_ptr(), q.stride(0), qbs, k.data_ptr(), k.stride(0), v.data_ptr(),
                                              v.stride(0), out.data_ptr(), out.stride(0), n_prompts, _stream()),
          "img2tok_attention")
    return out


def ln64_gelu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """In place on a contiguous bf16 buffer viewed as rows of 64: GELU(LayerNorm(row))."""
    _req_bf16(x, gamma, beta)
    assert x.is_contiguous() and x.numel() % 64 == 0
    check(_lib.lib().llmseg_ln64_gelu(x.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), x.numel() // 64,
                                      float(eps), _stream()), "ln64_gelu")
    return x


def mask_logits(up2: torch.Tensor, hyper: torch.Tensor, n_prompts: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """up2 bf16 [P*16384, 128], hyper bf16 [P,4,32] -> low-res logits fp32 [P,3,256,256]."""
    _req_bf16(up2, hyper)
    assert up2.is_contiguous() and hyper.is_contiguous() and up2.shape == (n_prompts * 16384, 128)
    if out is None:
        out = torch.empty((n_prompts, 3, 256, 256), dtype=torch.float32, device=up2.device)
    check(_lib.lib().llmseg_mask_logits(up2.data_ptr(), hyper.data_ptr(), n_prompts, out.data_ptr(), _stream()), "mask_logits")
    return out


def upscale_logits(up1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                   hyper: torch.Tensor, n_prompts: int, eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused LayerNorm2d(64)+GELU -> ConvTranspose #2 (+bias, GELU) -> hyper-network product: up1 bf16 [P*4096, 256],
    w2 bf16 [128,64], b2 bf16 [128], hyper bf16 [P,4,32] -> low-res logits fp32 [P,3,256,256]."""
    _req_bf16(up1, gamma, beta, w2, b2, hyper)
    assert up1.is_contiguous() and up1.shape == (n_prompts * 4096, 256) and w2.is_contiguous() and w2.shape == (128, 64)
    assert hyper.is_contiguous() and hyper.shape == (n_prompts, 4, 32) and b2.numel() == 128 and gamma.numel() == 64
    if out is None:
        out = torch.empty((n_prompts, 3, 256, 256), dtype=torch.float32, device=up1.device)
    check(_lib.lib().llmseg_upscale_logits(up1.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), w2.data_ptr(),
                                           b2.data_ptr(), hyper.data_ptr(), n_prompts, out.data_ptr(), _stream()),
          "upscale_logits")
    return out


def mask_stats(low_res: torch.Tensor, cand: Optional[torch.Tensor] = None, threshold: float = 0.0,
               offset: float = 1.0) -> torch.Tensor:
    """low_res fp32 [n,256,256] -> int32 [n,8] = {area, #(>t+o), #(>t-o), 1023-x0, 1023-y0, x1, y1, 0}."""
    _req_f32(low_res)
    _req_i32(cand)
    n = low_res.shape[0] if cand is None else cand.numel()
    stats = torch.empty((n, 8), dtype=torch.int32, device=low_res.device)
    check(_lib.lib().llmseg_mask_stats(low_res.data_ptr(), _ptr(cand), n, float(threshold), float(offset), stats.data_ptr(),
                                       _stream()), "mask_stats")
    return stats


def box_nms(boxes_sorted: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """boxes fp32 [n,4] XYXY sorted by score (descending) -> keep int32 [n]."""
    _req_f32(boxes_sorted)
    keep = torch.empty(boxes_sorted.shape[0], dtype=torch.int32, device=boxes_sorted.device)
    check(_lib.lib().llmseg_box_nms(boxes_sorted.data_ptr(), boxes_sorted.shape[0], float(iou_threshold), keep.data_ptr(),
                                    _stream()), "box_nms")
    return keep


def mask_soft(low_res: torch.Tensor, cand: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    """-> bf16 [n,256,256]: antialiased 1024 -> 256 resize of the binarised 4x up-sampled masks of the candidates."""
    _req_f32(low_res)
    _req_i32(cand)
    out = torch.empty((cand.numel(), 256, 256), dtype=torch.bfloat16, device=low_res.device)
    check(_lib.lib().llmseg_mask_soft(low_res.data_ptr(), cand.data_ptr(), cand.numel(), float(threshold), out.data_ptr(),
                                      _stream()), "mask_soft")
    return out


def mask_binarize(low_res: torch.Tensor, cand: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    """-> uint8 [n,1024,1024] binary masks of the candidates."""
    _req_f32(low_res)
    _req_i32(cand)
    out = torch.empty((cand.numel(), 1024, 1024), dtype=torch.uint8, device=low_res.device)
    check(_lib.lib().llmseg_mask_binarize(low_res.data_ptr(), cand.data_ptr(), cand.numel(), float(threshold),
                                          out.data_ptr(), _stream()), "mask_binarize")
    return out
