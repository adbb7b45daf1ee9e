"""Drop-in for tf_ops/3d_interpolation/tf_interpolate.py (reference :8-34): same names, same argument order.
The reference registers CPU kernels only (tf_interpolate.cpp:187,222,262); these run on the GPU."""
from . import _ops


def three_nn(xyz1, xyz2):
    """xyz1 (b,n,3) unknown points, xyz2 (b,m,3) known points -> dist (b,n,3) SQUARED distances ascending,
    idx (b,n,3) int32 (reference tf_interpolate.py:8-18).  No gradient."""
    return _ops.three_nn(xyz1, xyz2)


def three_interpolate(points, idx, weight):
    """points (b,m,c), idx (b,n,3) int32, weight (b,n,3) -> (b,n,c).  Differentiable w.r.t. points only
    (reference tf_interpolate.py:19-34)."""
    return _ops.ThreeInterpolate.apply(points, idx, weight)
