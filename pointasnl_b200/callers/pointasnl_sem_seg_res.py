"""Residual segmentation net -- caller-side mirror of the reference's models/pointasnl_sem_seg_res.py
(get_model :20-69, get_loss :72-88) on top of the drop-in layer API."""
import torch
import torch.nn.functional as F

from .. import _ops, tf_util
from ..pointasnl_util import PointASNLSetAbstraction, branch, get_repulsion_loss, index_cache, prefetch_neighbourhood
from ..pointnet_util import pointnet_fp_module
from .pointasnl_cls import l2_on_weights


def placeholder_inputs(batch_size, num_point, feature_channel=0, device="cuda"):
    c = 3 + feature_channel
    return (torch.zeros((batch_size, num_point, c), device=device),
            torch.zeros((batch_size, num_point), dtype=torch.int64, device=device),
            torch.zeros((batch_size, num_point), device=device))


def _sa(xyz, feat, npoint, mlp, scope, k_as, nl, cfg, prefetch_next=None):
    return PointASNLSetAbstraction(xyz, feat, npoint=npoint, nsample=32, mlp=mlp, scope=scope, as_neighbor=k_as,
                                   NL=nl, prefetch_next=prefetch_next, **cfg)


def get_model(point_cloud, is_training, num_class, bn_decay=None, weight_decay=None, feature_channel=0):
    """point_cloud (B, N, 3+feature_channel) -> per-point logits (B, N, num_class), end_points."""
    with index_cache():      # layer1_1 / layer1_2 repeat FPS + kNN on identical inputs (:35-36): computed once
        return _get_model(point_cloud, is_training, num_class, bn_decay, weight_decay, feature_channel)


def _get_model(point_cloud, is_training, num_class, bn_decay, weight_decay, feature_channel):
    end_points = {}
    num_point = point_cloud.shape[1]
    if feature_channel > 0 and point_cloud.is_cuda and point_cloud.shape[2] == 3 + feature_channel:
        l0_xyz, l0_points = _ops.split_cols(point_cloud, 3)
    elif feature_channel > 0:
        l0_xyz = point_cloud[:, :, 0:3].contiguous()
        l0_points = point_cloud[:, :, 3:3 + feature_channel].contiguous()
    else:
        l0_xyz = l0_points = point_cloud
    end_points["l0_xyz"] = l0_xyz
    cfg = dict(is_training=is_training, bn_decay=bn_decay, weight_decay=weight_decay)
    n1, n2, n3, n4 = num_point // 8, num_point // 32, num_point // 128, num_point // 256

    # layer1_1 / layer1_2's FPS + k-NN depend on l0_xyz only: started now, they run beside layer 0
    prefetch_neighbourhood(l0_xyz, n1, 32)
    _, l0_points = _sa(l0_xyz, l0_points, num_point, [16, 16, 32], "layer0", 0, False, cfg)
    # each residual stage: a sampling layer (FPS + AS + NL) and a same-resolution refinement layer, summed
    # layer1_2 reads the same inputs as layer1_1 and nothing of it: it runs beside it on a side stream
    res_branch = branch(l0_xyz, l0_points, slot=3)
    with res_branch:
        _, l1_2 = _sa(l0_xyz, l0_points, n1, [64, 64], "layer1_2", 0, False, cfg)
    l1_xyz, l1_1 = _sa(l0_xyz, l0_points, n1, [32, 32, 64], "layer1_1", 8, True, cfg, prefetch_next=(n2, 32))
    res_branch.join(l1_2)
    l1_2 = _ops.add_n(l1_2, l1_1)
    l2_xyz, l2_1 = _sa(l1_xyz, l1_2, n2, [64, 64, 128], "layer2_1", 4, True, cfg, prefetch_next=(n3, 32))
    _, l2_2 = _sa(l2_xyz, l2_1, n2, [128, 128], "layer2_2", 0, False, cfg)
    l2_2 = _ops.add_n(l2_2, l2_1)
    l3_xyz, l3_1 = _sa(l2_xyz, l2_2, n3, [128, 128, 256], "layer3_1", 0, True, cfg, prefetch_next=(n4, 32))
    _, l3_2 = _sa(l3_xyz, l3_1, n3, [256, 256], "layer3_2", 0, False, cfg)
    l3_2 = _ops.add_n(l3_2, l3_1)
    l4_xyz, l4_1 = _sa(l3_xyz, l3_1, n4, [256, 256, 512], "layer4_1", 0, True, cfg)   # sic: consumes l3_1 (:50)
    _, l4_2 = _sa(l4_xyz, l4_1, n4, [512, 512], "layer4_2", 0, False, cfg)
    l4_2 = _ops.add_n(l4_2, l4_1)
    end_points["l1_xyz"] = l1_xyz

    l3_points = pointnet_fp_module(l3_xyz, l4_xyz, l3_2, l4_2, [512, 512], is_training, bn_decay, scope="fa_layer1")
    l2_points = pointnet_fp_module(l2_xyz, l3_xyz, l2_2, l3_points, [256, 256], is_training, bn_decay, scope="fa_layer2")
    l1_points = pointnet_fp_module(l1_xyz, l2_xyz, l1_2, l2_points, [256, 128], is_training, bn_decay, scope="fa_layer3")
    l0_points = pointnet_fp_module(l0_xyz, l1_xyz, l0_points, l1_points, [128, 128, 128], is_training, bn_decay,
                                   scope="fa_layer4")

    net = tf_util.conv1d(l0_points, 128, 1, padding="VALID", activation_fn=tf_util.leaky_relu, bn=True,
                         is_training=is_training, scope="fc1", bn_decay=bn_decay, weight_decay=weight_decay)
    end_points["feats"] = net
    net = tf_util.dropout(net, keep_prob=0.5, is_training=is_training, scope="dp")
    net = tf_util.conv1d(net, num_class, 1, padding="VALID", activation_fn=None, weight_decay=weight_decay, scope="fc0")
    return net, end_points


def get_loss(pred, label, end_points, smpw=1.0, uniform_weight=0.01, weights_decay=1e-4, radius=0.07):
    """pred (B, N, C), label (B, N), smpw (B, N) -> total loss (:72-88).  tf.losses.sparse_softmax_cross_entropy with
    weights = weighted sum / number of non-zero weights."""
    store = tf_util.get_default_store()
    if pred.is_cuda and torch.is_tensor(smpw):
        classify_x2 = _ops.softmax_ce(pred, label, smpw, out_scale=2.0)    # see below: the reference counts it twice
    else:
        ce = F.cross_entropy(pred.reshape(-1, pred.shape[-1]), label.reshape(-1), reduction="none")
        if torch.is_tensor(smpw):
            w = smpw.reshape(-1)
            classify_loss = (ce * w).sum() / torch.count_nonzero(w).clamp(min=1)
        else:
            classify_loss = (ce * smpw).mean()
        classify_x2 = 2.0 * classify_loss
    uniform_loss = get_repulsion_loss(end_points["l1_xyz"], nsample=20, radius=radius)
    # :83 `weight_reg = tf.add_n(tf.get_collection('losses'))`.  tf.losses.sparse_softmax_cross_entropy registers its
    # result in that very collection (loss_collection=GraphKeys.LOSSES), next to the per-variable weight-decay terms
    # of tf_util._variable_with_weight_decay: weight_reg = classify_loss + sum(wd * l2_loss(var)), so the total (:85)
    # counts the classification loss TWICE.  (Every reference training script leaves weight_decay=None; without the
    # registered classify_loss `tf.add_n([])` would not even build.)  Reproduced, and pinned by
    # tests/golden/tf_graph_seg.npz, which executes the reference's own get_loss.
    total = classify_x2 + uniform_weight * uniform_loss
    if store.weight_decay:
        total = total + store.regularization_loss()
    if weights_decay:       # weights_decay=0: the caller folds the regulariser into its optimizer kernel (TrainStep)
        total = total + weights_decay * l2_on_weights(store)
    return total
