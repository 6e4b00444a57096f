import os
import sys

# CPU suite hygiene (must run before torch / libgomp initialise): forward runs on the main thread's OpenMP pool and
# backward on the autograd thread's own pool.  With the default of one pool thread per core both pools together
# oversubscribe the machine and their spin-waiting idle threads slowed single tests from 18 s to 8 min.  Four
# threads per pool and sleeping (not spinning) idle threads keep the whole suite at about a minute.
os.environ.setdefault("OMP_NUM_THREADS", "4")
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

import pytest  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
