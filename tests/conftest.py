"""pytest configuration: the ``gpu`` marker, repo-root imports and golden-fixture loading."""
import glob
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.pt"))):
        name = os.path.basename(p)[:-3]
        if kind is None or name.startswith(kind):
            out.append(name)
    return out


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def full_query(q_window, length):
    """Re-create a full-length Q whose last W rows are the stored window (the rest is never read)."""
    b, hq, w, d = q_window.shape
    q = torch.zeros(b, hq, length, d, dtype=q_window.dtype)
    q[:, :, length - w:, :] = q_window
    return q
