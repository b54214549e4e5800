import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arboris-python_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    """GPU tests are SKIPPED (not errors) on a machine without a CUDA device, so that a plain
    ``pytest tests`` works there too."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="needs a CUDA device")
        for it in gpu_items:
            it.add_marker(skip)


def load_golden(name):
    import numpy as np
    from arboris_b200.flatten import FlatModel
    model = FlatModel.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    with np.load(os.path.join(GOLDEN, "traj_%s.npz" % name)) as z:
        traj = {k: z[k] for k in z.files}
    return model, traj


@pytest.fixture(scope="session")
def golden():
    return load_golden
