"""ORACLE helper: import the reference's OWN unmodified Python files.

Works only where /root/reference exists (this build container, never the GPU
box).  Used by tests/golden/make_golden.py to generate the committed fixtures
and by the CPU tests that pin oracle/leco_ref.py against the real files.
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_DIR = os.environ.get("LECO_REFERENCE_DIR", "/root/reference")
_STUB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_stub")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "lora.py"))


def load_reference():
    """Returns a namespace with the reference modules lora, train_util,
    prompt_util, config_util, model_util imported from REFERENCE_DIR."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_DIR}")
    for p in (_STUB_DIR, REFERENCE_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    from types import SimpleNamespace
    mods = {}
    for name in ("lora", "model_util", "train_util", "prompt_util", "config_util"):
        mods[name] = importlib.import_module(name)
        origin = os.path.abspath(mods[name].__file__)
        assert origin.startswith(os.path.abspath(REFERENCE_DIR)), (name, origin)
    return SimpleNamespace(**mods)
