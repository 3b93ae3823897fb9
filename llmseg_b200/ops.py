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


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
         act: Optional[str] = None, residual: Optional[torch.Tensor] = None, res_mod: int = 0,
         out: Optional[torch.Tensor] = None, out_row_map: Optional[torch.Tensor] = None,
         out_rows: Optional[int] = None, swiglu: bool = False) -> torch.Tensor:
    """out = act(a @ w.T + bias) (+ residual).  a: [M,K] bf16, w: [N,K] bf16 (nn.Linear layout).

    out_row_map (int32 [M]) scatters GEMM row r to output row out_row_map[r] (negative = dropped);
    the residual is read at the same output row (modulo res_mod when given).
    swiglu: w rows are (gate0, up0, gate1, up1, ...) and out has N/2 columns.
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
    check(_lib.lib().llmseg_gemm(C.byref(p), _stream()), "gemm")
    return out


def gemm_qkv(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], q: torch.Tensor,
             k: torch.Tensor, vt: torch.Tensor, *, heads: int, head_dim: int, seq_in: int,
             seq_pad: int, rope_cos: Optional[torch.Tensor] = None,
             rope_sin: Optional[torch.Tensor] = None) -> None:
    """QKV projection writing q,k [(b*heads+h), seq_pad, hd] and vt [(b*heads+h), hd, seq_pad]."""
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
    check(_lib.lib().llmseg_gemm(C.byref(p), _stream()), "gemm_qkv")


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, *, batch: int,
              heads: int, head_dim: int, seq: int, seq_pad: int, scale: float, causal: bool = False,
              kv_len: Optional[torch.Tensor] = None, qext: Optional[torch.Tensor] = None,
              kext: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
              ext_cols: int = 0) -> torch.Tensor:
    """Fused attention over the q/k/vt buffers of gemm_qkv; out is [batch*seq, heads*head_dim]."""
    _req_bf16(q, k, vt, out, qext, kext, row_bias)
    p = AttnParams()
    p.q, p.k, p.vt = q.data_ptr(), k.data_ptr(), vt.data_ptr()
    p.out, p.ldo = out.data_ptr(), out.stride(0)
    p.batch, p.heads, p.head_dim, p.seq, p.seq_pad = batch, heads, head_dim, seq, seq_pad
    p.scale, p.causal = float(scale), int(causal)
    p.kv_len = _ptr(kv_len)
    p.ext_cols, p.qext, p.kext, p.row_bias = ext_cols, _ptr(qext), _ptr(kext), _ptr(row_bias)
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
    n_pad = (2 * t + 7) // 8 * 8
    out = torch.zeros(n_pad, rel_h.shape[1], dtype=torch.bfloat16, device=rel_h.device)
    out[:t] = rel_h
    out[t:2 * t] = rel_w
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


def rmsnorm(x: torch.Tensor, gamma: torch.Tensor, eps: float,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req_bf16(x, gamma, out)
    dim = x.shape[-1]
    x2 = x.reshape(-1, dim) if x.dim() != 2 else x
    if out is None:
        out = torch.empty_like(x2)
    check(_lib.lib().llmseg_rmsnorm(x2.data_ptr(), x2.stride(0), out.data_ptr(), out.stride(0),
                                    gamma.data_ptr(), x2.shape[0], dim, float(eps), _stream()),
          "rmsnorm")
    return out
