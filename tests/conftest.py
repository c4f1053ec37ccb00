import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (dev container only; skipped elsewhere)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/src/CyberBattleSim/cyberbattle")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    try:
        import torch

        have_gpu = torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "marlon_b200", "libcbx.so"))
    except Exception:
        have_gpu = False
    skip_gpu = pytest.mark.skip(reason="no CUDA device (or marlon_b200/libcbx.so not built): the GPU parity tests run on the B200 box")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)
