import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_built():
    """A fresh checkout has no binaries (they are git-ignored): build them once, exactly as
    __graft_entry__.build() does, so the suite never depends on the order the driver runs
    things in. nvcc cross-compiles sm_100a without a GPU."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "perturb_b200", "libperturb_gpu.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "perturb_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    if (not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libperturb_ref.so"))
            and os.path.isdir("/root/reference/src")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    _ensure_built()


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
