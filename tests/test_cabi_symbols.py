"""The C-ABI library loads on a CPU-only box and exports every symbol include/cellcomm_b200.h
declares (no compute calls here); the ctypes table covers exactly the header."""
import ctypes
import os
import re

from cellcomm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cellcomm_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 40
    raw = ctypes.CDLL(_lib.lib_path())
    for n in names:
        assert hasattr(raw, n), f"{n} declared in cellcomm_b200.h but not exported"
    assert lib.cc_arch() == b"sm_100a"
    assert lib.cc_version() >= 1
    assert lib.cc_launch_count() >= 0


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_gemm_desc_layout_matches_c_struct():
    """sizeof(cc_gemm_desc) as the C compiler lays it out (natural alignment, 4 segments)"""
    d = _lib.GemmDesc
    assert ctypes.sizeof(d) % 8 == 0
    assert d.a.offset % 8 == 0 and d.lda.offset == d.a.offset + 32
    assert d.k.size == 16


def test_errors_are_reported_not_swallowed():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.cc_mtx_load_csr(b"/nonexistent/file.mtx", ctypes.byref(h))
    assert rc != 0 and b"cannot open" in lib.cc_last_error()


def test_product_ops_refuse_cpu_tensors():
    """no CPU fallback: the op wrappers raise on host tensors instead of computing there"""
    import pytest
    import torch
    from cellcomm_b200 import ops
    a = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ops.gemm(4, 4, [a], [a], [64], 0, 0, out16=a)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ops.copy2d(a, a)
