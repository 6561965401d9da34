import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_close_to_max(a, b, rel=1e-5, what=""):
    """max|a-b| <= rel * max|b| over the whole array (per-cell relative error is meaningless
    for near-cancelling +- cells); NaN patterns must agree."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert (na == nb).all(), what + ": NaN pattern differs"
    if nb.all():
        return
    scale = np.abs(b[~nb]).max()
    err = np.abs(a[~nb] - b[~nb]).max() if (~nb).any() else 0.0
    assert err <= rel * max(scale, 1e-30), "%s: max err %.3e vs scale %.3e (rel %.2e > %.1e)" % (
        what, err, scale, err / max(scale, 1e-30), rel)


@pytest.fixture(scope="session")
def oracle():
    from oracle import evk_oracle
    evk_oracle.build()
    return evk_oracle


def make_events(seed, n, H, W, frac=True, tscale=0.1, pol="pm1", dtype=np.float32):
    rng = np.random.default_rng(seed)
    xs = rng.random(n) * (W - 1) if frac else rng.integers(0, W, n).astype(np.float64)
    ys = rng.random(n) * (H - 1) if frac else rng.integers(0, H, n).astype(np.float64)
    ts = np.sort(rng.random(n)) * tscale
    if pol == "pm1":
        ps = rng.integers(0, 2, n) * 2.0 - 1.0
    elif pol == "ones":
        ps = np.ones(n)
    else:
        ps = rng.standard_normal(n)
    return tuple(a.astype(dtype) for a in (xs, ys, ts, ps))
