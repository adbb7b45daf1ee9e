"""Drop-in for utils/pointasnl_util.py (reference): same function names, positional order and return arity.

Point-set functions (this file, part 1): knn_query, sampling, grouping  (reference :22-76).
Layer functions (part 2, below): SampleWeights, AdaptiveSampling, PointNonLocalCell, PointASNLSetAbstraction,
PointASNLDecodingLayer, get_repulsion_loss -- added with the TF-style variable store in tf_util.
"""
import torch

from . import _ops, tf_grouping, tf_sampling  # noqa: F401


def knn_query(k, support_pts, query_pts):
    """support_pts (B, N1, 3), query_pts (B, N2, 3) -> neighbor_idx (B, N2, k) int32 (reference :22-30).
    The reference leaves the device for a nanoflann kd-tree behind tf.py_func; this is one CUDA kernel with the same
    exact-kNN contract (ascending squared distance, the query itself first when it is a support point)."""
    return _ops.knn(k, support_pts, query_pts)


def sampling(npoint, pts, feature=None):
    """pts (B, N, D) -> sub_pts (B, npoint, D) [, sub_feature (B, npoint, C)] by farthest point sampling on pts
    (reference :33-49).  Differentiable w.r.t. pts / feature (tf.gather_nd semantics)."""
    fps_idx = tf_sampling.farthest_point_sample(npoint, pts.detach())
    sub = _ops.GatherRows.apply(pts, fps_idx)
    if feature is None:
        return sub
    return sub, _ops.GatherRows.apply(feature, fps_idx)


def grouping(feature, K, src_xyz, q_xyz, use_xyz=True, use_knn=True, radius=0.2):
    """K neighbours of every q_xyz in src_xyz (reference :51-76) -> grouped_xyz (B, npoint, K, 3),
    grouped_feature (B, npoint, K, [3+]C), idx.

    use_knn=True is the path every shipped model takes.  use_knn=False is BROKEN in the reference (it uses `idx`
    before defining it, :67-71); here it is given the obviously intended meaning (ball query + group_point).
    idx is returned as the (B, npoint, K) int32 neighbour index (the reference returns the (B, npoint, K, 2)
    [batch, point] pairs tf.gather_nd needs; no caller reads it)."""
    if use_knn:
        idx = knn_query(K, src_xyz.detach(), q_xyz.detach())
    else:
        idx, _ = tf_grouping.query_ball_point(radius, K, src_xyz.detach(), q_xyz.detach())
    if use_xyz:
        grouped_feature = _ops.GroupConcat.apply(src_xyz, feature, idx, None)   # [xyz | feature], one pass
        grouped_xyz = grouped_feature[..., :3]
    else:
        grouped_xyz = tf_grouping.group_point(src_xyz, idx)
        grouped_feature = tf_grouping.group_point(feature, idx)
    return grouped_xyz, grouped_feature, idx
