"""Deterministic, torch-version-independent weight fill shared by the golden generator and the tests.

TEST INFRASTRUCTURE ONLY.  Every tensor of a state dict is drawn from numpy's legacy RandomState in key
order, scaled by role, so a fixture only has to store a seed and the (key, shape) schema instead of
megabytes of weights.
"""
import numpy as np
import torch


def fill_state_dict(schema, seed):
    """schema: list of (key, shape).  Returns {key: fp32 tensor}."""
    rs = np.random.RandomState(seed)
    out = {}
    for key, shape in schema:
        shape = tuple(shape)
        a = rs.standard_normal(shape).astype(np.float64)
        leaf = key.rsplit(".", 1)[-1]
        if leaf in ("weight_u", "weight_v"):
            a = a / max(np.linalg.norm(a), 1e-12)
        elif leaf == "gamma":
            a = 0.5 + 0.1 * a                      # non-zero so attention contributes
        elif leaf in ("weight_orig", "weight") and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            a = a * (1.0 / np.sqrt(fan_in))
        elif leaf == "weight" and len(shape) == 1:  # InstanceNorm affine scale
            a = 1.0 + 0.1 * a
        elif leaf == "bias":
            a = 0.1 * a
        out[key] = torch.from_numpy(a.astype(np.float32)).clone()
    return out


def seeded(shape, seed, scale=1.0, clamp=None):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal(tuple(shape)).astype(np.float32) * scale
    if clamp is not None:
        a = np.clip(a, -clamp, clamp)
    return torch.from_numpy(a).clone()


def schema_of(module):
    return [(k, list(v.shape)) for k, v in module.state_dict().items()]
