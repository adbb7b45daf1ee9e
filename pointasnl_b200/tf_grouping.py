"""Drop-in for tf_ops/grouping/tf_grouping.py (reference :8-73): same names, same argument order."""
import torch

from . import _ops


def query_ball_point(radius, nsample, xyz1, xyz2):
    """radius float, nsample int, xyz1 (b, ndataset, 3) support, xyz2 (b, npoint, 3) queries ->
    idx (b, npoint, nsample) int32, pts_cnt (b, npoint) int32 (reference tf_grouping.py:8-21).  No gradient."""
    return _ops.ball_query(radius, nsample, xyz1, xyz2)


def group_point(points, idx):
    """points (b, ndataset, c), idx (b, npoint, nsample) int32 -> (b, npoint, nsample, c).
    Differentiable w.r.t. points (reference tf_grouping.py:33-46)."""
    return _ops.GroupPoint.apply(points, idx)


def knn_point(k, xyz1, xyz2):
    """k int, xyz1 (b, ndataset, 3) support, xyz2 (b, npoint, 3) queries -> val (b, npoint, k) squared L2
    distances ascending, idx (b, npoint, k) int32 (reference tf_grouping.py:48-73).  No gradient.
    The reference materialises the (b, npoint, ndataset) distance matrix and selection-sorts it; here it is one
    tiled kernel: same distances position by position; equidistant neighbours come in ascending index order
    (the reference's swap-based selection sort leaves them in a swap-history-dependent order)."""
    return _ops.knn(k, xyz1, xyz2, return_dist=True)


def select_top_k(k, dist):
    """dist (b, m, n) -> idx (b, m, n) int32, dist_out (b, m, n): the first k entries of each row are the k smallest in
    ascending order; the other n-k entries are the permuted remainder (reference tf_grouping.py:22-32 ->
    selection_sort_gpu, tf_grouping_g.cu:83-123).  Same "first strict minimum, then swap" rule as the reference, so
    equal distances and the remainder come out exactly where the reference leaves them.  No gradient."""
    return _ops.selection_sort(k, dist)
