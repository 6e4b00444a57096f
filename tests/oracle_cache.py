"""Oracle results used by the GPU tests are deterministic functions of seeds, so they are computed ONCE
in the build container (fp32 CPU, minutes) and committed under tests/golden/cache/; on the GPU box the
tests load them instead of burning GPU-box minutes on CPU work.  A missing entry is recomputed on the
spot (same code path), so the cache is an optimisation, never a different answer.

    python tests/oracle_cache.py        # (re)generate every entry
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
CACHE_DIR = os.path.join(ROOT, "tests", "golden", "cache")


def cached(name: str, fn, write: bool = False):
    path = os.path.join(CACHE_DIR, name + ".pt")
    if os.path.exists(path) and not write:
        return torch.load(path)
    torch.set_num_threads(min(8, os.cpu_count() or 8) if "tiny" in name else (os.cpu_count() or 8))
    val = fn()
    if write or os.environ.get("LECO_WRITE_FIXTURES") == "1":
        os.makedirs(CACHE_DIR, exist_ok=True)
        torch.save(val, path)
    return val


def main():
    from tests.gpu_checks import kernel_cases as kc
    for arch, n, hw in (("tiny21", 2, 16), ("tiny15", 2, 16), ("tinyxl", 2, 16), ("tiny21", 2, 32), ("sd21", 2, 64),
                        ("sd15", 2, 64), ("sdxl", 2, 128)):
        cached(f"fwd_{arch}_{n}_{hw}", lambda: kc.oracle_forward(arch, n, hw), write=True)
        print("fwd", arch, n, hw, flush=True)
    for arch in ("tiny21", "tiny15"):
        cached(f"grads_{arch}", lambda: kc.oracle_grads(arch, 2, 8, None), write=True)
        print("grads", arch, flush=True)
    cached("grads_c3lier_tiny15", lambda: kc.oracle_grads("tiny15", 2, 8, "c3lier"), write=True)
    print("grads c3lier", flush=True)
    cached("grads_rank32_tiny21", lambda: kc.oracle_grads("tiny21", 2, 8, "rank32"), write=True)
    print("grads rank32", flush=True)
    from __graft_entry__ import oracle_iterations
    cached("iters_tiny21", lambda: oracle_iterations(3), write=True)
    print("iters", flush=True)
    from __graft_entry__ import _SETTINGS_DYN
    cached("iters_tiny21_dyn", lambda: oracle_iterations(4, settings=_SETTINGS_DYN), write=True)
    print("iters dyn", flush=True)
    cached("iters_tiny21_multi", lambda: oracle_iterations(5, multi=True), write=True)
    print("iters multi", flush=True)
    cached("iters_tiny21_lms", lambda: oracle_iterations(3, scheduler="lms"), write=True)
    print("iters lms", flush=True)
    from __graft_entry__ import oracle_iterations_xl
    cached("iters_tinyxl", lambda: oracle_iterations_xl(3), write=True)
    print("iters xl", flush=True)


if __name__ == "__main__":
    main()
