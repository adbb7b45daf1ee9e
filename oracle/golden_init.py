"""Name-seeded variable values shared by the golden generators and every implementation under test
(TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import zlib

import numpy as np


def golden_init(name, shape, kind):
    """Deterministic variable values shared by every implementation under test: seeded by the variable NAME, so that
    a golden file needs to store no weights.  Weights: uniform in the xavier range; BN gamma in [0.5, 1.5], beta and
    biases in [-0.2, 0.2] (non-trivial values exercise every term); moving averages at TF's initial values."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    shape = tuple(int(s) for s in shape)
    if kind == "weights":
        fan_in = int(np.prod(shape[:-1]))
        fan_out = int(np.prod(shape[:-2])) * shape[-1] if len(shape) > 2 else shape[-1]
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-limit, limit, shape)
    if kind == "gamma":
        return rng.uniform(0.5, 1.5, shape)
    if kind in ("beta", "biases"):
        return rng.uniform(-0.2, 0.2, shape)
    if kind == "moving_mean":
        return np.zeros(shape)
    if kind == "moving_variance":
        return np.ones(shape)
    raise KeyError(kind)


def digest(name, g):
    """[sum, l2 norm, <g, r(name)>] of a tensor (float64): three numbers that pin a gradient without storing it."""
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    r = np.random.default_rng(zlib.crc32(("dir:" + name).encode())).standard_normal(g.size)
    return np.array([g.sum(), np.sqrt((g * g).sum()), (g * r).sum()])


def probe_weights(name, shape):
    """Name-seeded random weights of a scalar probe sum(t * w): makes every output element matter in the gradient."""
    return np.random.default_rng(zlib.crc32(("probe:" + name).encode())).standard_normal(tuple(int(s) for s in shape))
