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


class Routing:
    """Record the encoders' global-max arg-max positions in one evaluation and force them in others (oracle or CUDA path), keyed by
    (encoder name, call index) so the order in which the two-stream path happens to build its branches does not matter.  Isolates the
    one discontinuous operator of the path: with the routing pinned, every remaining operation is smooth in its operands."""

    def __init__(self):
        self.idx = {}

    # ---- oracle side: install as oracle.gim_oracle.GMAX_HOOK ----
    def oracle_hook(self, net, record):
        counts = {}

        def hook(prefix, x):
            key = "%s.%s" % (net, prefix)
            j = counts.get(key, 0)
            counts[key] = j + 1
            flat = x.flatten(2)
            if record:
                self.idx[(key, j)] = flat.argmax(-1)
            return flat.gather(2, self.idx[(key, j)].to(flat.device).unsqueeze(-1)).squeeze(-1)
        return hook

    # ---- CUDA path: patches gim_img_models.Encoder.forward for the modules given as {module: name} ----
    def patch_encoders(self, named_modules, record=False):
        import contextlib

        import torch
        from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
        from optimalstrategiesagainstgenerativeattacks_b200 import model_blocks as mb
        from optimalstrategiesagainstgenerativeattacks_b200 import ops
        names = {id(m): n for m, n in named_modules.items()}
        counts = {}
        routing = self

        def forward(self, x):
            x = M._as_nhwc(x)
            mb.sn_prepare_module(self, skip=() if self.att_loc < self.n_down_blocks else (self.att,))
            for i, block in enumerate(self.down_blocks):
                if i == self.att_loc:
                    x = self.att(x)
                x = block(x, want_ops=i + 1 < self.n_down_blocks)
            if isinstance(x, ops.Act):
                x = x.t32
            key = names[id(self)]
            j = counts.get(key, 0)
            counts[key] = j + 1
            if record:
                n, h, w, c = x.shape
                routing.idx[(key, j)] = x.detach().reshape(n, h * w, c).argmax(1)
            idx = routing.idx[(key, j)].to(device=x.device, dtype=torch.int32).contiguous()
            x = ops.GatherIdxFn.apply(x, idx)
            return ops.lrelu(x) if self.use_out_lrelu else x

        @contextlib.contextmanager
        def ctx():
            orig = M.Encoder.forward
            M.Encoder.forward = forward
            try:
                yield
            finally:
                M.Encoder.forward = orig
        return ctx()
