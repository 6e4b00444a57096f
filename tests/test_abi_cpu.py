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


def test_header_is_plain_c_and_a_c_host_can_bind_it(tmp_path):
    """The boundary is a C ABI (plain pointers and sizes, no C++ / torch types in the signatures): `include/leco_b200.h`
    compiles as C99, and a C host that includes it, dlopens the library and calls through the declared prototypes gets
    the version and the error convention (negative code + leco_last_error() text) without a GPU."""
    import shutil
    import subprocess
    import pytest
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "leco_b200.h")], check=True)
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "leco_b200.h"
typedef int (*abi_fn)(void);
typedef const char* (*err_fn)(void);
typedef int (*opt_fn)(float*, void*, float*, float*, float*, const void*, const float*, int64_t, int, void*);
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  abi_fn abi = (abi_fn)dlsym(h, "leco_abi_version");
  err_fn err = (err_fn)dlsym(h, "leco_last_error");
  opt_fn opt = (opt_fn)dlsym(h, "leco_optim_flat_master");
  opt_fn same = leco_optim_flat_master;   /* the header's prototype has exactly this type */
  (void)same;
  if (!abi || !err || !opt) return 3;
  int rc = opt(NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, 1, NULL);
  printf("%d %d %s\n", abi(), rc, err());
  return (abi() == 2 && rc < 0 && strstr(err(), "leco_optim_flat_master") != NULL) ? 0 : 1;
}
''')
    exe = tmp_path / "host"
    # the prototype reference `same = leco_optim_flat_master` needs the symbol at link time: link against the library
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", f"-I{inc}", str(src), "-o", str(exe), "-ldl",
                    capi.lib_path(), f"-Wl,-rpath,{os.path.dirname(capi.lib_path())}"], check=True)
    out = subprocess.run([str(exe), capi.lib_path()], capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
