import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build everything in-tree once per session (no-op when up to date)."""
    from libecp_b200 import build

    build.build_product()
    build.build_hostcheck()
    build.build_oracle()


def load_matrix(name):
    z = np.load(os.path.join(GOLDEN, f"{name}_matrix.npz"))
    dim = int(z["dim"])
    M = np.zeros((dim, dim))
    M[np.triu_indices(dim)] = z["triu"]
    return M


def load_blocks(name):
    z = np.load(os.path.join(GOLDEN, f"{name}_blocks.npz"))
    return z["keys"], z["off"], z["vals"]


def assert_parity(got, ref, what=""):
    """north_star tolerance: |x - ref| <= 1e-12 + 1e-10 |ref| element-wise."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, what
    assert not np.isnan(got).any(), what + ": NaN in result"
    err = np.abs(got - ref)
    tol = 1e-12 + 1e-10 * np.abs(ref)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} elements out of tolerance, max |d| = {err.max():.3e}"
