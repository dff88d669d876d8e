import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
BOX = 1000.0


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a CPU-only machine skips the GPU tests instead of erroring in their fixtures."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ma_golden():
    return dict(np.load(os.path.join(GOLDEN, "ma_golden.npz")))


@pytest.fixture(scope="session")
def pk_golden():
    return dict(np.load(os.path.join(GOLDEN, "pk_golden.npz")))


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/cpu.py over oracle/liboracle.so). Test infrastructure only."""
    from oracle import build as obuild
    obuild.build()
    from oracle import cpu
    return cpu


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor), NaNs (empty 2D bins: 0/0) must coincide."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs"
    if a.size == 0 or na.all():
        return 0.0
    den = np.maximum(np.abs(b), floor)
    den = np.where(den == 0, 1.0, den)
    return float(np.nanmax(np.abs(a - b) / den))


def make_particles(seed, n, clustered=False, box=BOX, sigma=0.04):
    """Uniform particles; `clustered` puts half of them into 16 Gaussian clumps of width sigma*box
    (sigma=0.04 -> density contrast of a few tens; sigma=0.01 -> thousands of particles per cell)."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32) * np.float32(box)
    if clustered:
        c = rng.random((16, 3), dtype=np.float32) * np.float32(box)
        k = n // 2
        pos[:k] = (c[rng.integers(0, 16, k)] + rng.normal(0, box * sigma, (k, 3)).astype(np.float32)) % np.float32(box)
    edge = np.array([[0, 0, 0], [box, box, box], [np.nextafter(np.float32(box), np.float32(0)), 0.5, box / 2],
                     [box, 0, np.nextafter(np.float32(box), np.float32(0))]], dtype=np.float32)
    pos[:len(edge)] = edge
    W = rng.random(n, dtype=np.float32)
    return pos, W
