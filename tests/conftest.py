import json
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


@pytest.fixture(scope="session")
def schemas():
    with open(os.path.join(GOLDEN, "schemas.json")) as f:
        return json.load(f)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().double().cpu().numpy()
    return np.asarray(a, dtype=np.float64).ravel()


def rel_err(a, b):
    """||a-b|| / ||b|| in float64 (the 'rel' of north_star's parity gates)."""
    a = _np(a)
    b = _np(b)
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


N_SAMPLES = 128


def grad_sample_index(numel, j):
    """Element positions stored per parameter tensor by oracle/make_golden.py (`sample_index`)."""
    if numel <= N_SAMPLES:
        return np.arange(numel)
    return np.sort(np.random.RandomState(7919 + j).choice(numel, N_SAMPLES, replace=False))


def sample_rows(tensors):
    """[n, N_SAMPLES] float64 samples of a list of tensors (None -> NaN row), same layout as the goldens' `*_gsamp`."""
    rows = np.full((len(tensors), N_SAMPLES), np.nan)
    for j, t in enumerate(tensors):
        if t is None:
            continue
        a = t.detach().double().reshape(-1).cpu().numpy()
        idx = grad_sample_index(a.size, j)
        rows[j, :idx.size] = a[idx]
    return rows


def sample_errs(got, want, zero_floor=1e-9):
    """Per-tensor relative error ||got - want|| / ||want|| over the stored ELEMENTS of each tensor.  Rows whose true gradient is
    (numerically) identically zero relative to the largest tensor are returned as NaN: they carry pure rounding noise in every
    implementation (conv biases feeding Instance/AdaIN norms, the mapper bias under mean removal)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    none_w, none_g = np.isnan(want).all(1), np.isnan(got).all(1)
    assert (none_w == none_g).all(), "set of parameters that receive gradients differs"
    w = np.nan_to_num(want)
    g = np.nan_to_num(got)
    wn = np.linalg.norm(w, axis=1)
    err = np.linalg.norm(g - w, axis=1) / np.maximum(wn, 1e-300)
    err[none_w | (wn < zero_floor * wn.max())] = np.nan
    return err
