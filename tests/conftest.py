import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a host without CUDA skips the gpu-marked tests instead of failing them."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a B200 (no CUDA device on this host)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names(prefix=None, exclude_prefix=()):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if prefix is not None:
        names = [n for n in names if n.startswith(prefix)]
    if prefix is None:  # agent_* records hold one training epoch of the reference agent, not a column run
        exclude_prefix = tuple(exclude_prefix) + ("agent_",)
    return [n for n in names if not any(n.startswith(e) for e in exclude_prefix)]


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


# Tolerance of the north star: 1e-9 relative with an absolute floor of 1e-12 cm.
RTOL, ATOL = 1e-9, 1e-12


def max_excess(a, b, rtol=RTOL, atol=ATOL):
    """max over elements of |a-b| / (atol + rtol*|b|); <= 1 means within tolerance."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    # the reference can carry NaN through a step before one of its guards raises (golden nan_dry_depth_col185:
    # torch.min propagates NaN): NaN in both = equal, NaN in one = infinitely wrong
    na, nb = np.isnan(a), np.isnan(b)
    if (na != nb).any():
        return float("inf")
    both = na & nb
    if both.all():
        return 0.0
    a, b = a[~both], b[~both]
    return float(np.max(np.abs(a - b) / (atol + rtol * np.abs(b))))
