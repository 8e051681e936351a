import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def oracle_ds(ds):
    """atlite_b200 Dataset -> plain dict the oracle consumes."""
    d = {k: np.asarray(ds.raw(k)) for k in ds.keys()}
    d["time"] = ds.coords["time"]
    d["lon"] = np.asarray(ds.coords["lon"])
    d["lat"] = np.asarray(ds.coords["lat"])
    return d


def assert_parity(got, want, capacity=None, rtol=1e-4, atol_cap=1e-6, what=""):
    """The parity bar of this repo: fp32 kernels vs the float64 oracle,
    |gpu - oracle| <= 1e-4 * |oracle| + 1e-6 * capacity_bus   (capacity = row sum of
    the aggregation matrix; 1 for per-cell / per-unit outputs)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    cap = 1.0 if capacity is None else np.asarray(capacity, dtype=np.float64)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN positions differ ({nan_g.sum()} vs {nan_w.sum()})"
    tol = rtol * np.abs(want) + atol_cap * np.maximum(cap, 1e-30)
    err = np.abs(got - want)
    bad = (err > tol) & ~nan_w
    if bad.any():
        i = np.unravel_index(np.nanargmax(np.where(bad, err / np.maximum(tol, 1e-300), 0)), got.shape)
        raise AssertionError(
            f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst at {i}: "
            f"got {got[i]!r} want {want[i]!r} (err {err[i]:.3e}, tol {np.broadcast_to(tol, got.shape)[i]:.3e})"
        )


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False
