"""ModelNet40 classifier -- caller-side mirror of the reference's models/pointasnl_cls.py (get_model :17-52,
get_loss :55-70) on top of the drop-in layer API.  Variables are created in the active tf_util.VariableStore under
the reference's scope names (layer1/..., layer2/..., layer3_1, layer3_2, fc1, fc2, fc3)."""
import torch
import torch.nn.functional as F

from .. import _ops, tf_util
from ..pointasnl_util import PointASNLSetAbstraction, branch, get_repulsion_loss, index_cache
from ..pointnet_util import pointnet_sa_module

# (scope, npoint, nsample, mlp) of the two set-abstraction layers, reference :33-36
SA_LAYERS = (("layer1", 512, 32, [64, 64, 128]), ("layer2", 128, 64, [128, 128, 256]))
# global heads on (l1_xyz, l1_points) and (l2_xyz, l2_points), reference :39-40
HEADS = (("layer3_1", [128, 256, 512]), ("layer3_2", [256, 512, 1024]))


def placeholder_inputs(batch_size, num_point, use_normal=False, device="cuda"):
    """Shapes/dtypes of the reference's placeholders (:11-15), as zero tensors."""
    c = 6 if use_normal else 3
    return (torch.zeros((batch_size, num_point, c), device=device),
            torch.zeros((batch_size,), dtype=torch.int64, device=device))


def get_model(point_cloud, is_training, use_normal=False, bn_decay=None, weight_decay=None, num_class=40,
              adaptive_sample=False):
    """point_cloud (B, N, 3 or 6) -> logits (B, num_class), end_points."""
    with index_cache():
        return _get_model(point_cloud, is_training, use_normal, bn_decay, weight_decay, num_class, adaptive_sample)


def _get_model(point_cloud, is_training, use_normal, bn_decay, weight_decay, num_class, adaptive_sample):
    batch_size = point_cloud.shape[0]
    end_points = {}
    if use_normal and point_cloud.is_cuda:
        xyz, points = _ops.split_cols(point_cloud, 3)
    elif use_normal:
        xyz, points = point_cloud[:, :, 0:3].contiguous(), point_cloud[:, :, 3:6].contiguous()
    else:
        xyz = points = point_cloud
    end_points["l0_xyz"] = xyz
    as_neighbor = [12, 12] if adaptive_sample else [0, 0]

    levels = []
    pooled = [None, None]
    head_branch = None
    for li, ((scope, npoint, nsample, mlp), k_as) in enumerate(zip(SA_LAYERS, as_neighbor)):
        nxt = (SA_LAYERS[li + 1][1], SA_LAYERS[li + 1][2]) if li + 1 < len(SA_LAYERS) else None
        xyz, points = PointASNLSetAbstraction(xyz, points, npoint=npoint, nsample=nsample, mlp=mlp,
                                              is_training=is_training, bn_decay=bn_decay, weight_decay=weight_decay,
                                              scope=scope, as_neighbor=k_as, prefetch_next=nxt)
        levels.append((xyz, points))
        if li == 0:
            # the global head on (l1_xyz, l1_points) (:39) only needs layer 1: it runs beside layer 2 on a side stream
            head_branch = branch(xyz, points, slot=2)
            with head_branch:
                hs, hm = HEADS[0]
                _, feat, _ = pointnet_sa_module(xyz, points, npoint=None, radius=None, nsample=None, mlp=hm, mlp2=None,
                                                group_all=True, is_training=is_training, bn_decay=bn_decay, scope=hs)
                pooled[0] = feat.reshape(batch_size, -1)
    end_points["l1_xyz"] = levels[0][0]
    end_points["l2_xyz"] = levels[0][0]   # sic: the reference stores l1_xyz under 'l2_xyz' (:38)

    (scope, mlp), (lx, lp) = HEADS[1], levels[1]
    _, feat, _ = pointnet_sa_module(lx, lp, npoint=None, radius=None, nsample=None, mlp=mlp, mlp2=None,
                                    group_all=True, is_training=is_training, bn_decay=bn_decay, scope=scope)
    pooled[1] = feat.reshape(batch_size, -1)
    head_branch.join(pooled[0])
    net = _ops.concat_cols(pooled[1], pooled[0])          # [l3_points, l3_points_res] (:45)
    net = tf_util.fully_connected(net, 512, bn=True, is_training=is_training, scope="fc1", bn_decay=bn_decay)
    net = tf_util.dropout(net, keep_prob=0.4, is_training=is_training, scope="dp1")
    net = tf_util.fully_connected(net, 256, bn=True, is_training=is_training, scope="fc2", bn_decay=bn_decay)
    net = tf_util.dropout(net, keep_prob=0.4, is_training=is_training, scope="dp2")
    net = tf_util.fully_connected(net, num_class, activation_fn=None, scope="fc3")
    return net, end_points


def l2_on_weights(store):
    """sum of tf.nn.l2_loss(v) = sum(v^2)/2 over every variable whose name contains 'weights' (:58)."""
    return tf_util.l2_loss_sum([v for name, v in store.variables().items() if "weights" in name])


def get_loss(pred, label, end_points, uniform_weight=0, weights_decay=1e-4):
    """pred (B, num_class), label (B,) -> total loss (:55-70)."""
    store = tf_util.get_default_store()
    classify_loss = _ops.softmax_ce(pred, label) if pred.is_cuda else F.cross_entropy(pred, label)
    total = classify_loss
    if uniform_weight > 0:
        total = total + uniform_weight * get_repulsion_loss(end_points["l1_xyz"], nsample=20, radius=0.07)
    if weights_decay:       # weights_decay=0: the caller folds the regulariser into its optimizer kernel (TrainStep)
        total = total + weights_decay * l2_on_weights(store)
    return total
