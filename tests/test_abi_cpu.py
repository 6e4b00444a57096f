"""CPU: the C-ABI library builds for sm_100a, dlopens without a GPU and exports every symbol
include/leco_b200.h declares.  No compute call is made."""
import os
import re

from leco_b200 import capi
from leco_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    path = build_library(verbose=False)
    assert os.path.exists(path)
    lib = capi.load()
    assert lib.leco_abi_version() == 2
    assert lib.leco_launch_count() >= 0


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "leco_b200.h")).read()
    declared = set(re.findall(r"\b(leco_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"leco_gemm_args"}
    lib = capi.load()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(capi.EXPORTED), declared ^ set(capi.EXPORTED)


def test_product_path_has_no_cpu_fallback():
    import pytest
    import torch
    from leco_b200 import ops
    a = torch.zeros((8, 16), dtype=torch.bfloat16)
    with pytest.raises(capi.LecoError):
        ops.gemm(a, a)
    # nothing under leco_b200/ may import the oracle
    pkg = os.path.join(ROOT, "leco_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_argument_errors_come_back_as_codes_not_crashes():
    """Error convention of the boundary: no exception / abort crosses the C ABI; a bad call returns a
    negative code and leaves the reason in leco_last_error().  These checks run before any CUDA call, so they work on
    a GPU-less host."""
    import ctypes
    lib = capi.load()
    lib.leco_last_error.restype = ctypes.c_char_p
    g = capi.GemmArgs()
    assert lib.leco_gemm_bf16(ctypes.byref(g), None) < 0
    assert b"null operand" in lib.leco_last_error()
    dummy = ctypes.c_void_p(0x1000)                      # never dereferenced: validation fails first
    g.a, g.b, g.d = dummy, dummy, dummy
    g.M, g.N, g.K, g.lda, g.ldb, g.ldd = 128, 100, 64, 64, 64, 104
    assert lib.leco_gemm_bf16(ctypes.byref(g), None) < 0
    assert b"N%8" in lib.leco_last_error()
    g.N, g.ldd, g.mode, g.cc, g.cn, g.ch, g.cw = 64, 64, 1, 32, 1, 8, 16   # conv with C % 64 != 0
    g.K = 9 * 32
    g.ldb = g.K
    assert lib.leco_gemm_bf16(ctypes.byref(g), None) < 0
    assert b"conv needs" in lib.leco_last_error()
    import pytest
    with pytest.raises(capi.LecoError):
        capi.check(-1, "leco_gemm_bf16")
