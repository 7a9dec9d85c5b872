import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    """-> dict of torch tensors (0-dim arrays become python floats/ints)."""
    out = {}
    with np.load(os.path.join(GOLDEN, name + '.npz')) as f:
        for k in f.files:
            a = f[k]
            out[k] = a.item() if a.ndim == 0 else torch.from_numpy(a.copy())
    return out


def tables_of(g):
    return (g['centroids'], g['matrices'], float(g['temperature']), float(g['regularization']))


def rel_fro(a, b):
    """per-matrix relative Frobenius error, max over the batch."""
    a = a.double().reshape(a.shape[0], -1)
    b = b.double().reshape(b.shape[0], -1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300)).max().item()


@pytest.fixture
def golden():
    return load_golden
