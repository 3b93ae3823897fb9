"""ctypes binding of libllmseg_b200.so (the C ABI in include/llmseg_b200.h).

There is deliberately NO fallback: if the shared object is missing, or a call returns a negative
LLMSEG_E* code (e.g. LLMSEG_EARCH on a non-sm_100 device), a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os as _os

# LLMSEG_B200_LIB: load another build of the same library (A/B measurements of kernel changes on one box)
LIB_PATH = Path(_os.environ["LLMSEG_B200_LIB"]) if _os.environ.get("LLMSEG_B200_LIB") else _PKG / "libllmseg_b200.so"

c_void_p, c_int, c_float = C.c_void_p, C.c_int, C.c_float


class GemmParams(C.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_int),
        ("W", c_void_p), ("ldw", c_int),
        ("C", c_void_p), ("ldc", c_int),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_int), ("res_mod", c_int),
        ("act", c_int), ("mode", c_int),
        ("out_row_map", c_void_p),
        ("q", c_void_p), ("k", c_void_p), ("vt", c_void_p),
        ("heads", c_int), ("head_dim", c_int), ("seq_in", c_int), ("seq_pad", c_int),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", C.c_size_t),
        ("row_stats", c_void_p),
        ("row_stats_parts", c_int), ("norm_dim", c_int), ("norm_eps", c_float), ("norm_rms", c_int),
        ("stats_out", c_void_p),
        ("stats_final", c_void_p), ("stats_dim", c_int), ("stats_eps", c_float), ("stats_rms", c_int),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("vt", c_void_p),
        ("out", c_void_p), ("ldo", c_int),
        ("batch", c_int), ("heads", c_int), ("head_dim", c_int), ("seq", c_int), ("seq_pad", c_int),
        ("scale", c_float), ("causal", c_int),
        ("kv_len", c_void_p),
        ("ext_cols", c_int), ("qext", c_void_p), ("kext", c_void_p), ("row_bias", c_void_p),
        ("out_row_map", c_void_p),
    ]


_lib = None


def _declare(lib):
    lib.llmseg_last_error.restype = C.c_char_p
    lib.llmseg_version.restype = c_int
    lib.llmseg_launch_count.restype = C.c_uint64
    for name in SYMBOLS:
        getattr(lib, name)  # raises AttributeError if the .so is stale
    lib.llmseg_gemm.argtypes = [C.POINTER(GemmParams), c_void_p]
    lib.llmseg_gemm_workspace_bytes.restype = C.c_size_t
    lib.llmseg_gemm_stats_parts.argtypes = [c_int, c_int]
    lib.llmseg_attention.argtypes = [C.POINTER(AttnParams), c_void_p]
    lib.llmseg_relpos_prep.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_float, c_void_p, c_int, c_void_p, c_void_p]
    lib.llmseg_layernorm.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                     c_int, c_float, c_void_p, c_void_p]
    lib.llmseg_norm_stats.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]
    lib.llmseg_rmsnorm.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                   c_float, c_void_p, c_void_p]
    lib.llmseg_patchify.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.llmseg_embed_splice.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int, c_int, c_int, c_int, C.c_int64, C.c_int64,
                                        c_int, c_void_p]
    lib.llmseg_add_rows_bcast.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_void_p]
    lib.llmseg_gather_rows.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.llmseg_fill_kv_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p]
    lib.llmseg_im2col3x3.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.llmseg_maskpool_workspace.argtypes = [c_int]
    lib.llmseg_maskpool_workspace.restype = C.c_size_t
    lib.llmseg_maskpool.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    lib.llmseg_small_attention.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                           c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                           c_int, c_void_p]
    lib.llmseg_select.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.llmseg_align_iou_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                          c_void_p]
    lib.llmseg_selector_losses.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                           c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]
    lib.llmseg_lm_cross_entropy.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                            C.c_int64, C.c_int64, c_void_p, c_void_p, c_void_p]
    lib.llmseg_dice_bce_loss.argtypes = [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p,
                                         c_void_p, c_void_p]
    ll = C.c_longlong
    lib.llmseg_point_tokens.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]
    lib.llmseg_tok2img_attention.argtypes = [c_void_p, c_int, c_void_p, c_int, ll, c_void_p, c_int, ll, c_void_p, c_int,
                                             c_int, c_void_p]
    lib.llmseg_img2tok_attention.argtypes = [c_void_p, c_int, ll, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                             c_int, c_void_p]
    lib.llmseg_ln64_gelu.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, ll, c_float, c_void_p]
    lib.llmseg_mask_logits.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.llmseg_upscale_logits.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                          c_void_p]
    lib.llmseg_mask_stats.argtypes = [c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p]
    lib.llmseg_box_nms.argtypes = [c_void_p, c_int, c_float, c_void_p, c_void_p]
    lib.llmseg_mask_soft.argtypes = [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]
    lib.llmseg_mask_binarize.argtypes = [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]


# every symbol include/llmseg_b200.h declares (tests/test_abi.py checks header <-> .so <-> this list)
SYMBOLS = [
    "llmseg_last_error", "llmseg_version", "llmseg_launch_count",
    "llmseg_gemm", "llmseg_gemm_workspace_bytes", "llmseg_gemm_stats_parts", "llmseg_attention", "llmseg_relpos_prep", "llmseg_layernorm", "llmseg_rmsnorm", "llmseg_norm_stats",
    "llmseg_patchify", "llmseg_embed_splice", "llmseg_add_rows_bcast", "llmseg_gather_rows", "llmseg_fill_kv_rows", "llmseg_im2col3x3",
    "llmseg_maskpool_workspace", "llmseg_maskpool", "llmseg_small_attention", "llmseg_select",
    "llmseg_align_iou_loss", "llmseg_dice_bce_loss", "llmseg_selector_losses", "llmseg_lm_cross_entropy",
    "llmseg_point_tokens", "llmseg_tok2img_attention", "llmseg_img2tok_attention", "llmseg_ln64_gelu", "llmseg_mask_logits",
    "llmseg_upscale_logits",
    "llmseg_mask_stats", "llmseg_box_nms", "llmseg_mask_soft", "llmseg_mask_binarize",
]


def lib():
    """Load (once) and return the CDLL; raise loudly when the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m llmseg_b200.build` "
                "(llmseg_b200 has no CPU or PyTorch fallback path)")
        l = C.CDLL(str(LIB_PATH))
        _declare(l)
        _lib = l
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().llmseg_last_error().decode(errors="replace")
        raise RuntimeError(f"llmseg_b200 {what} failed ({code}): {msg}")


def launch_count() -> int:
    return int(lib().llmseg_launch_count())
