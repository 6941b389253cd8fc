import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pyoracle as po  # noqa: E402  (test infrastructure)

GOLDEN = os.path.join(ROOT, "tests", "golden")
TERRAINS = ("rough_terrain", "slope", "synth_nan", "synth_mixed")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    po.build(ref=True)


@pytest.fixture(scope="session", params=TERRAINS)
def golden(request):
    name = request.param
    T = po.Terrain.from_npz(os.path.join(GOLDEN, f"terrain_{name}.npz"))
    G = dict(np.load(os.path.join(GOLDEN, f"golden_{name}.npz")))
    return name, T, G


def load_terrain(name):
    return po.Terrain.from_npz(os.path.join(GOLDEN, f"terrain_{name}.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bits_equal(a, b, where=None, what=""):
    """Bit-exact fp64 comparison; entries where the reference left NaN (unwritten) are skipped."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    m = ~np.isnan(b)
    if where is not None:
        m &= where.reshape(where.shape + (1,) * (b.ndim - where.ndim)).astype(bool)
    bad = (bits(a) != bits(b)) & m & ~((a == 0) & (b == 0))  # +0 / -0 compare equal
    assert not bad.any(), f"{what}: {int(bad.sum())} of {int(m.sum())} fp64 values differ in bits"
