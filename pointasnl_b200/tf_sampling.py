"""Drop-in for tf_ops/sampling/tf_sampling.py (reference :13-57): same names, same argument order."""
from . import _ops


def farthest_point_sample(npoint, inp):
    """npoint: int; inp (batch_size, ndataset, 3) float32 -> (batch_size, npoint) int32.  No gradient
    (reference tf_sampling.py:48-57).  Note the argument order quirk: the Python shim takes (npoint, inp)."""
    return _ops.fps(npoint, inp)


def gather_point(inp, idx):
    """inp (batch_size, ndataset, 3), idx (batch_size, npoints) int32 -> (batch_size, npoints, 3).
    Differentiable w.r.t. inp (reference tf_sampling.py:29-47)."""
    return _ops.GatherPoint.apply(inp, idx)


def prob_sample(inp, inpr):
    """inp (batch_size, ncategory) float32 non-negative weights, inpr (batch_size, npoints) float32 uniform numbers ->
    (batch_size, npoints) int32: categories drawn with probability proportional to inp (reference tf_sampling.py:13-22 ->
    cumsumKernel + binarysearchKernel, tf_sampling_g.cu:7-104).  Same fp32 summation order, hence the same indices."""
    return _ops.prob_sample(inp, inpr)
