"""CPU: the C-ABI shared object loads and exports every symbol include/llmseg_b200.h declares
(no compute calls without a GPU), and the product path refuses to run without CUDA."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / "include" / "llmseg_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(llmseg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from llmseg_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = _header_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in llmseg_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms, "llmseg_b200/_lib.py SYMBOLS out of sync with the header"
    lib.llmseg_version.restype = ctypes.c_int
    assert lib.llmseg_version() >= 100


def test_struct_layouts_match_header_order():
    """ctypes Structure field order must follow the C structs (same names, same order)."""
    from llmseg_b200 import _lib
    text = (ROOT / "include" / "llmseg_b200.h").read_text()
    for cname, struct in (("llmseg_gemm_params", _lib.GemmParams), ("llmseg_attn_params", _lib.AttnParams)):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname, text).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
        assert names == [f[0] for f in struct._fields_], cname


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_ops_fail_loudly_without_gpu():
    from llmseg_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, a)
