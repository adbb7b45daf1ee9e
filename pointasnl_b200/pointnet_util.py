"""Drop-in for the two functions of utils/pointnet_util.py the PointASNL models call (reference
models/pointasnl_cls.py:39-40, models/pointasnl_sem_seg_res.py:57-60): pointnet_sa_module with group_all=True and
pointnet_fp_module.  The multi-scale / radius-grouping variants of PointNet++ are unused by the shipped models
 and are not mirrored."""
import torch

from . import _ops, tf_util
from .tf_interpolate import three_interpolate, three_nn


def sample_and_group_all(xyz, points, use_xyz=True):
    """xyz (B, N, 3), points (B, N, C) -> new_xyz (B, 1, 3) zeros, new_points (B, 1, N, 3+C), idx, grouped_xyz
    (reference pointnet_util.py:57-82)."""
    batch_size, nsample, _ = xyz.shape
    new_xyz = torch.zeros((batch_size, 1, 3), device=xyz.device, dtype=xyz.dtype)
    idx = torch.arange(nsample, device=xyz.device, dtype=torch.int32).reshape(1, 1, nsample).repeat(batch_size, 1, 1)
    grouped_xyz = xyz.reshape(batch_size, 1, nsample, 3)
    if points is not None:
        new_points = _ops.concat_cols(xyz, points) if use_xyz else points
        new_points = new_points.unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, group_all, is_training, bn_decay, scope,
                       bn=True, pooling="max", knn=False, use_xyz=True, use_nchw=False):
    """PointNet set-abstraction module (reference pointnet_util.py:85-150); only group_all=True with max pooling is
    reachable from the PointASNL models."""
    if not group_all or pooling != "max" or use_nchw:
        raise NotImplementedError("only group_all=True, pooling='max', NHWC is used by the PointASNL models")
    with tf_util.variable_scope(scope):
        new_xyz, new_points, idx, _ = sample_and_group_all(xyz, points, use_xyz)
        for i, num_out_channel in enumerate(mlp):
            new_points = tf_util.conv2d(new_points, num_out_channel, [1, 1], padding="VALID", stride=[1, 1], bn=bn,
                                        is_training=is_training, scope="conv%d" % i, bn_decay=bn_decay)
        new_points = (_ops.MaxPoolK.apply(new_points) if new_points.is_cuda else new_points.amax(dim=2)).unsqueeze(2)
        if mlp2 is not None:
            for i, num_out_channel in enumerate(mlp2):
                new_points = tf_util.conv2d(new_points, num_out_channel, [1, 1], padding="VALID", stride=[1, 1],
                                            bn=bn, is_training=is_training, scope="conv_post_%d" % i,
                                            bn_decay=bn_decay)
        return new_xyz, new_points.squeeze(2), idx


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope, bn=True):
    """Feature propagation: inverse-squared-distance interpolation of points2 (on xyz2, sparse) onto xyz1 (dense),
    concat with points1, 1x1 conv stack (reference pointnet_util.py:199-229).  three_nn / three_interpolate run on
    the GPU here (CPU-only, single-threaded TF ops in the reference)."""
    with tf_util.variable_scope(scope):
        dist, idx = three_nn(xyz1.detach(), xyz2.detach())
        dist = torch.clamp(dist, min=1e-10)
        norm = (1.0 / dist).sum(dim=2, keepdim=True)
        weight = (1.0 / dist) / norm
        interpolated_points = three_interpolate(points2, idx, weight)
        if points1 is not None:
            new_points1 = _ops.concat_cols(interpolated_points, points1)
        else:
            new_points1 = interpolated_points
        new_points1 = new_points1.unsqueeze(2)
        for i, num_out_channel in enumerate(mlp):
            new_points1 = tf_util.conv2d(new_points1, num_out_channel, [1, 1], padding="VALID", stride=[1, 1], bn=bn,
                                         is_training=is_training, scope="conv_%d" % i, bn_decay=bn_decay)
        return new_points1.squeeze(2)
