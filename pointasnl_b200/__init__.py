"""pointasnl_b200 -- the PointASNL per-layer point-set hot path as hand-written sm_100a CUDA kernels.

Drop-in modules carrying the reference's own names and argument orders (yanx27/PointASNL):

    tf_sampling     farthest_point_sample, gather_point                      (tf_ops/sampling/tf_sampling.py)
    tf_grouping     query_ball_point, group_point, knn_point, select_top_k   (tf_ops/grouping/tf_grouping.py)
    tf_interpolate  three_nn, three_interpolate                              (tf_ops/3d_interpolation/tf_interpolate.py)
    pointasnl_util  knn_query, sampling, grouping, AdaptiveSampling, PointNonLocalCell,
                    PointASNLSetAbstraction, PointASNLDecodingLayer, get_repulsion_loss   (utils/pointasnl_util.py)
    pointnet_util   pointnet_sa_module (group_all), pointnet_fp_module       (utils/pointnet_util.py)
    tf_util         conv1d / conv2d / fully_connected / batch_norm / dropout on a TF-style variable store

All of them take and return ``torch.Tensor`` on a CUDA device; the compute is libpasnl_b200.so (include/pasnl.h),
reached through ctypes.  There is no CPU path: importing works anywhere, calling an op without the built library
or with CPU tensors raises.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def library_version():
    return _lib.version()
