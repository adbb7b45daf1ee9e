"""Layer- and model-level CPU restatement of the reference's TF graph code (TEST INFRASTRUCTURE, see
oracle/__init__.py).  **Pinned**: tests/test_oracle.py checks every function here against
tests/golden/tf_graph_*.npz, which were produced by executing the reference's own Python files (imported unmodified
from /root/reference) on the stand-in tensorflow module tests/golden/tf_shim.py -- outputs to 1e-9, gradient digests of
every variable to 1e-7, fp64.

It follows utils/pointasnl_util.py, utils/pointnet_util.py, utils/tf_util.py and the two model files line by line,
in the reference's own (unfused) formulation: gather_nd-style indexing, materialised attention maps, separate
softmax / matmul / reduce ops.  Native pieces run the reference's own CPU code when oracle/_ref has it
(nanoflann kNN, three_nn / three_interpolate) and the C restatement (oracle/cpu_ref.c) otherwise (FPS, ball query).

Two uses:
  * parity: given the product's variables (``store.variables()``), ``cls_forward`` / ``seg_forward`` must
    reproduce the CUDA path's logits, loss and parameter gradients within tolerance (tests/test_gpu_model.py);
  * CPU baseline: ``bench.py --impl reference`` / the ``cpu_baseline`` leg time a forward+backward of this
    restatement on the host cores.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cref, refnative


# ---------------------------------------------------------------------------------------------- native ops
def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def fps_idx(npoint, xyz):
    return torch.from_numpy(cref.fps(_np(xyz), npoint)).long()


def knn_idx(k, support, query, use_reference=True):
    """The kd-tree the reference calls through tf.py_func (pointasnl_util.py:22-30) when oracle/_ref has it."""
    if use_reference and refnative.available("ref_knn.so"):
        return torch.from_numpy(refnative.knn_batch(_np(support), _np(query), k)).long()
    return torch.from_numpy(cref.knn(k, _np(support), _np(query))).long()


def ball_idx(radius, nsample, xyz1, xyz2):
    idx, cnt = cref.ball_query(radius, nsample, _np(xyz1), _np(xyz2))
    return torch.from_numpy(idx).long(), torch.from_numpy(cnt)


def three_nn(xyz1, xyz2):
    fn = refnative.three_nn if refnative.available("ref_interp.so") else cref.three_nn
    d, i = fn(_np(xyz1), _np(xyz2))
    return torch.from_numpy(d).to(xyz1.dtype), torch.from_numpy(i).long()


class _ThreeInterpolateRef(torch.autograd.Function):
    """The reference's CPU op pair (tf_interpolate.cpp:107-153), fp32, used for the timed CPU baseline."""

    @staticmethod
    def forward(ctx, points, idx, weight):
        fn = refnative.three_interpolate if refnative.available("ref_interp.so") else cref.three_interpolate
        out = fn(_np(points), idx.numpy().astype(np.int32), _np(weight))
        ctx.save_for_backward(idx, weight)
        ctx.m = points.shape[1]
        return torch.from_numpy(out).to(points.dtype)

    @staticmethod
    def backward(ctx, g):
        idx, weight = ctx.saved_tensors
        fn = refnative.three_interpolate_grad if refnative.available("ref_interp.so") else cref.three_interpolate_grad
        out = fn(ctx.m, idx.numpy().astype(np.int32), _np(weight), _np(g))
        return torch.from_numpy(out).to(g.dtype), None, None


def three_interpolate(points, idx, weight, native=False):
    if native:
        return _ThreeInterpolateRef.apply(points, idx, weight)
    b = torch.arange(points.shape[0])[:, None, None]
    return (points[b, idx] * weight.unsqueeze(-1)).sum(dim=2)      # same maths, any dtype, autograd-native


def gather_nd(t, idx):
    """tf.gather_nd(t, [batch, idx]) for idx (B, ...) of point indices."""
    b = torch.arange(t.shape[0]).reshape((-1,) + (1,) * (idx.dim() - 1))
    return t[b, idx]


# ---------------------------------------------------------------------------------------------- dense glue
class Ref:
    """Holds the variables (TF name -> tensor) and evaluation switches."""

    def __init__(self, variables=None, dtype=torch.float32, dropout=True, native_knn=True, native_interp=False, seed=0,
                 init="tf"):
        """variables: TF name -> tensor (e.g. the product's store.variables()), or None to create them lazily: init="tf"
        = the reference's initialisers (xavier-uniform weights, zero biases, BN gamma 1 / beta 0; used by the CPU
        baseline), init="golden" = oracle.golden_init.golden_init(name, shape, kind), the name-seeded values the
        tests/golden/tf_graph_*.npz vectors were generated with."""
        self.lazy = variables is None
        self.init = init
        self.v = {} if self.lazy else {
            k: v.detach().cpu().to(dtype).requires_grad_(v.dtype.is_floating_point and "moving_" not in k)
            for k, v in variables.items()}
        self.dtype = dtype
        self.dropout = dropout
        self.native_knn = native_knn
        self.native_interp = native_interp
        self.scope = []
        self.gen = torch.Generator().manual_seed(seed)

    def var(self, name, shape=None, kind=None):
        full = "/".join(self.scope + [name])
        if full not in self.v:
            if not self.lazy:
                raise KeyError(full)
            if self.init == "golden":
                from .golden_init import golden_init
                t = torch.from_numpy(golden_init(full, shape, name))
            elif kind == "xavier":      # tf.contrib.layers.xavier_initializer (uniform), tf_util.py:_variable_with_weight_decay
                fan_in = int(np.prod(shape[:-1]))
                fan_out = int(np.prod(shape[:-2])) * shape[-1] if len(shape) > 2 else shape[-1]
                limit = math.sqrt(6.0 / (fan_in + fan_out))
                t = (torch.rand(shape, generator=self.gen) * 2 - 1) * limit
            else:
                t = torch.full(shape, 1.0 if kind == "ones" else 0.0)
            self.v[full] = t.to(self.dtype).requires_grad_(True)
        return self.v[full]

    class _Scope:
        def __init__(self, ref, name):
            self.ref, self.name = ref, name

        def __enter__(self):
            self.ref.scope.append(self.name)

        def __exit__(self, *a):
            self.ref.scope.pop()

    def vs(self, name):
        return Ref._Scope(self, name)

    # tf_util.batch_norm_template (:512-531): batch statistics (biased variance), eps 1e-3
    def bn(self, x, scope):
        c = x.shape[-1]
        with self.vs(scope):
            beta, gamma = self.var("beta", (c,), "zeros"), self.var("gamma", (c,), "ones")
        dims = tuple(range(x.dim() - 1))
        mean = x.mean(dim=dims, keepdim=True)
        var = x.var(dim=dims, unbiased=False, keepdim=True)
        return (x - mean) / torch.sqrt(var + 1e-3) * gamma + beta

    # tf_util.conv1d / conv2d / fully_connected (:60-115, :120-185, :318-355): matmul + bias -> bn -> activation
    def dense(self, x, scope, cout, bn=False, act=F.relu, full_width=False):
        with self.vs(scope):
            cin = x.shape[-1]
            wshape = (1, x.shape[2], cin, cout) if full_width else ((1, 1, cin, cout) if x.dim() == 4 else
                                                                     ((1, cin, cout) if x.dim() == 3 else (cin, cout)))
            w = self.var("weights", wshape, "xavier")
            b = self.var("biases", (cout,), "zeros")
            if full_width:                      # conv2d kernel [1, W], VALID: contraction over (W, Cin)
                out = torch.matmul(x.reshape(x.shape[0], x.shape[1], 1, -1), w.reshape(-1, w.shape[-1])) + b
            else:
                out = torch.matmul(x, w.reshape(-1, w.shape[-1])) + b
            if bn:
                out = self.bn(out, "bn")
        return act(out) if act is not None else out

    def drop(self, x, keep_prob):
        if not self.dropout:
            return x
        mask = (torch.rand(x.shape, generator=self.gen) < keep_prob).to(x.dtype)
        return x * mask / keep_prob


# ---------------------------------------------------------------------------------------------- pointasnl_util.py
def sampling(r, npoint, pts, feature):
    """:33-49"""
    idx = fps_idx(npoint, pts)
    return gather_nd(pts, idx), gather_nd(feature, idx)


def grouping(r, feature, K, src_xyz, q_xyz, use_xyz=True):
    """:51-76 (use_knn=True branch)"""
    idx = knn_idx(K, src_xyz, q_xyz, r.native_knn)
    grouped_xyz = gather_nd(src_xyz, idx)
    grouped_feature = gather_nd(feature, idx)
    if use_xyz:
        grouped_feature = torch.cat([grouped_xyz, grouped_feature], dim=-1)
    return grouped_xyz, grouped_feature, idx


def sample_weights(r, new_point, grouped_xyz, mlps, scope, bn=True, scaled=True):
    """:112-156"""
    with r.vs(scope):
        channel = new_point.shape[-1]
        bottleneck = max(32, channel // 2)
        normalized_xyz = grouped_xyz - grouped_xyz[:, :, 0:1, :]
        new_point = torch.cat([normalized_xyz, new_point], dim=-1)
        transformed_feature = r.dense(new_point, "conv_kv_ds", 2 * bottleneck, bn=bn, act=None)
        transformed_new_point = r.dense(new_point, "conv_query_ds", bottleneck, bn=bn, act=None)
        transformed_feature1 = transformed_feature[..., :bottleneck]
        feature = transformed_feature[..., bottleneck:]
        weights = torch.matmul(transformed_new_point, transformed_feature1.transpose(-1, -2))
        if scaled:
            weights = weights / math.sqrt(float(bottleneck))
        weights = torch.softmax(weights, dim=-1)
        net = torch.matmul(weights, feature)
        for i, c in enumerate(mlps):
            net = r.dense(net, "mlp2_%d" % i, c, bn=bn, act=F.relu if i < len(mlps) - 1 else None)
        return torch.softmax(net, dim=2)


def adaptive_sampling(r, group_xyz, group_feature, num_neighbor, scope, bn=True):
    """:158-173"""
    with r.vs(scope):
        num_channel = group_feature.shape[-1]
        if num_neighbor == 0:
            return group_xyz[:, :, 0, :], group_feature[:, :, 0, :]
        shift_group_xyz = group_xyz[:, :, :num_neighbor, :]
        shift_group_points = group_feature[:, :, :num_neighbor, :]
        sample_weight = sample_weights(r, shift_group_points, shift_group_xyz, [32, 1 + num_channel], scope, bn)
        new_weight_xyz = sample_weight[..., 0:1].expand(-1, -1, -1, 3)
        new_weight_feature = sample_weight[..., 1:]
        new_xyz = (shift_group_xyz * new_weight_xyz).sum(dim=2)
        new_feature = (shift_group_points * new_weight_feature).sum(dim=2)
        return new_xyz, new_feature


def point_nonlocal_cell(r, feature, new_point, mlp, scope, bn=True, scaled=True):
    """:175-219, mode='dot'; the (B, npoint*nsample, ndataset) attention map is materialised as in the reference."""
    with r.vs(scope):
        bottleneck = mlp[0]
        batch_size, npoint, nsample, _ = new_point.shape
        transformed_feature = r.dense(feature.unsqueeze(2), "conv_kv", 2 * bottleneck, bn=bn, act=None)
        transformed_new_point = r.dense(new_point, "conv_query", bottleneck, bn=bn, act=None)
        transformed_new_point = transformed_new_point.reshape(batch_size, npoint * nsample, bottleneck)
        transformed_feature1 = transformed_feature[..., :bottleneck].squeeze(2)
        transformed_feature2 = transformed_feature[..., bottleneck:].squeeze(2)
        attention_map = torch.matmul(transformed_new_point, transformed_feature1.transpose(-1, -2))
        if scaled:
            attention_map = attention_map / math.sqrt(float(bottleneck))
        attention_map = torch.softmax(attention_map, dim=-1)
        new_nonlocal_point = torch.matmul(attention_map, transformed_feature2)
        new_nonlocal_point = r.dense(new_nonlocal_point.reshape(batch_size, npoint, nsample, bottleneck),
                                     "conv_back_project", mlp[-1], bn=bn)
        return new_nonlocal_point.squeeze(1)


def set_abstraction(r, xyz, feature, npoint, nsample, mlp, scope, bn=True, as_neighbor=8, NL=True):
    """:221-292"""
    with r.vs(scope):
        num_points, num_channel = feature.shape[1], feature.shape[2]
        if num_points == npoint:
            new_xyz, new_feature = xyz, feature
        else:
            new_xyz, new_feature = sampling(r, npoint, xyz, feature)
        grouped_xyz, new_point, _ = grouping(r, feature, nsample, xyz, new_xyz)
        nl_channel = mlp[-1]
        if num_points != npoint:
            new_xyz, new_feature = adaptive_sampling(r, grouped_xyz, new_point, as_neighbor, scope, bn)
        grouped_xyz = grouped_xyz - new_xyz.unsqueeze(2)
        new_point = torch.cat([grouped_xyz, new_point], dim=-1)
        if NL:
            new_nonlocal_point = point_nonlocal_cell(r, feature, new_feature.unsqueeze(1),
                                                     [max(32, num_channel // 2), nl_channel], scope, bn)
        skip_spatial = new_point.amax(dim=2)
        skip_spatial = r.dense(skip_spatial, "skip", mlp[-1], bn=bn)
        for i in range(len(mlp) - 1):
            new_point = r.dense(new_point, "conv%d" % i, mlp[i], bn=bn)
        with r.vs("weight_net"):
            weight = r.dense(grouped_xyz, "wconv0", 32, bn=True)
        new_point = torch.matmul(new_point.transpose(2, 3), weight)
        new_point = r.dense(new_point, "after_conv", mlp[-1], bn=bn, full_width=True).squeeze(2)
        new_point = new_point + skip_spatial
        if NL:
            new_point = new_point + new_nonlocal_point
        new_point = r.dense(new_point, "aggregation", mlp[-1], bn=bn)
        return new_xyz, new_point


def get_repulsion_loss(r, pred, nsample=20, radius=0.07):
    """:361-378"""
    idx, _ = ball_idx(radius, nsample, pred, pred)
    grouped_pred = gather_nd(pred, idx) - pred.unsqueeze(2)
    h = 0.03
    dist_square = (grouped_pred ** 2).sum(dim=-1)
    neg, _ = torch.topk(-dist_square, 5, dim=-1)
    dist_square = torch.clamp(-neg[:, :, 1:], min=1e-12)
    dist = torch.sqrt(dist_square)
    weight = torch.exp(-dist_square / h ** 2)
    return torch.mean(radius - dist * weight)


def decoding_layer(r, xyz1, xyz2, points1, points2, nsample, mlp, scope, bn=True):
    """PointASNLDecodingLayer (:294-351), NL=False (the default every caller uses)"""
    with r.vs(scope):
        dist, idx = three_nn(xyz1, xyz2)
        dist = torch.clamp(dist, min=1e-10)
        norm = (1.0 / dist).sum(dim=2, keepdim=True)
        weight = (1.0 / dist) / norm
        interpolated = three_interpolate(points2, idx, weight, r.native_interp)
        grouped_xyz, grouped_feature, _ = grouping(r, interpolated, nsample, xyz1, xyz1)
        grouped_xyz = grouped_xyz - xyz1.unsqueeze(2)
        with r.vs("decode_weight_net"):
            w = r.dense(grouped_xyz, "wconv0", 32, bn=True)
        new_points = torch.matmul(grouped_feature.transpose(2, 3), w)
        new_points = r.dense(new_points, "decode_after_conv", mlp[0], bn=bn, full_width=True)
        new_points1 = torch.cat([new_points, points1.unsqueeze(2)], dim=-1) if points1 is not None else new_points
        for i in range(1, len(mlp)):
            new_points1 = r.dense(new_points1, "conv_%d" % i, mlp[i], bn=bn)
        return new_points1.squeeze(2)


# ---------------------------------------------------------------------------------------------- pointnet_util.py
def sa_group_all(r, xyz, points, mlp, scope):
    """pointnet_sa_module(group_all=True) (pointnet_util.py:85-150 with :57-82)"""
    with r.vs(scope):
        new_points = torch.cat([xyz, points], dim=2).unsqueeze(1)
        for i in range(len(mlp)):
            new_points = r.dense(new_points, "conv%d" % i, mlp[i], bn=True)
        return new_points.amax(dim=2)


def fp_module(r, xyz1, xyz2, points1, points2, mlp, scope):
    """pointnet_fp_module (pointnet_util.py:199-229)"""
    with r.vs(scope):
        dist, idx = three_nn(xyz1, xyz2)
        dist = torch.clamp(dist, min=1e-10)
        norm = (1.0 / dist).sum(dim=2, keepdim=True)
        weight = (1.0 / dist) / norm
        interpolated = three_interpolate(points2, idx, weight, r.native_interp)
        new_points1 = torch.cat([interpolated, points1], dim=2) if points1 is not None else interpolated
        new_points1 = new_points1.unsqueeze(2)
        for i in range(len(mlp)):
            new_points1 = r.dense(new_points1, "conv_%d" % i, mlp[i], bn=True)
        return new_points1.squeeze(2)


# ---------------------------------------------------------------------------------------------- models
def l2_on_weights(r):
    return torch.stack([0.5 * (v ** 2).sum() for k, v in r.v.items() if "weights" in k]).sum()


def cls_forward(r, point_cloud, label=None, use_normal=True, adaptive_sample=False, uniform_weight=0):
    """models/pointasnl_cls.py get_model (:17-52) + get_loss (:55-70)"""
    pc = point_cloud.to(r.dtype)
    batch_size = pc.shape[0]
    if use_normal:
        l0_xyz, l0_points = pc[:, :, 0:3], pc[:, :, 3:6]
    else:
        l0_xyz = l0_points = pc
    as_neighbor = [12, 12] if adaptive_sample else [0, 0]
    l1_xyz, l1_points = set_abstraction(r, l0_xyz, l0_points, 512, 32, [64, 64, 128], "layer1", as_neighbor=as_neighbor[0])
    l2_xyz, l2_points = set_abstraction(r, l1_xyz, l1_points, 128, 64, [128, 128, 256], "layer2", as_neighbor=as_neighbor[1])
    l3_points_res = sa_group_all(r, l1_xyz, l1_points, [128, 256, 512], "layer3_1")
    l3_points = sa_group_all(r, l2_xyz, l2_points, [256, 512, 1024], "layer3_2")
    net = torch.cat([l3_points.reshape(batch_size, -1), l3_points_res.reshape(batch_size, -1)], dim=-1)
    net = r.drop(r.dense(net, "fc1", 512, bn=True), 0.4)
    net = r.drop(r.dense(net, "fc2", 256, bn=True), 0.4)
    net = r.dense(net, "fc3", 40, act=None)
    if label is None:
        return net, None
    classify_loss = F.cross_entropy(net, label)
    uniform_loss = get_repulsion_loss(r, l1_xyz) if uniform_weight > 0 else classify_loss
    total = classify_loss + uniform_weight * uniform_loss + 1e-4 * l2_on_weights(r)
    return net, total


def seg_forward(r, point_cloud, num_class, label=None, smpw=None, feature_channel=3, uniform_weight=0.01, radius=0.07):
    """models/pointasnl_sem_seg_res.py get_model (:20-69) + get_loss (:72-88)"""
    pc = point_cloud.to(r.dtype)
    num_point = pc.shape[1]
    if feature_channel > 0:
        l0_xyz, l0_points = pc[:, :, 0:3], pc[:, :, 3:3 + feature_channel]
    else:
        l0_xyz = l0_points = pc
    n = [num_point // 8, num_point // 32, num_point // 128, num_point // 256]
    sa = set_abstraction
    _, l0_points = sa(r, l0_xyz, l0_points, num_point, 32, [16, 16, 32], "layer0", as_neighbor=0, NL=False)
    l1_xyz, l1_1 = sa(r, l0_xyz, l0_points, n[0], 32, [32, 32, 64], "layer1_1", as_neighbor=8)
    _, l1_2 = sa(r, l0_xyz, l0_points, n[0], 32, [64, 64], "layer1_2", as_neighbor=0, NL=False)
    l1_2 = l1_2 + l1_1
    l2_xyz, l2_1 = sa(r, l1_xyz, l1_2, n[1], 32, [64, 64, 128], "layer2_1", as_neighbor=4)
    _, l2_2 = sa(r, l2_xyz, l2_1, n[1], 32, [128, 128], "layer2_2", as_neighbor=0, NL=False)
    l2_2 = l2_2 + l2_1
    l3_xyz, l3_1 = sa(r, l2_xyz, l2_2, n[2], 32, [128, 128, 256], "layer3_1", as_neighbor=0)
    _, l3_2 = sa(r, l3_xyz, l3_1, n[2], 32, [256, 256], "layer3_2", as_neighbor=0, NL=False)
    l3_2 = l3_2 + l3_1
    l4_xyz, l4_1 = sa(r, l3_xyz, l3_1, n[3], 32, [256, 256, 512], "layer4_1", as_neighbor=0)
    _, l4_2 = sa(r, l4_xyz, l4_1, n[3], 32, [512, 512], "layer4_2", as_neighbor=0, NL=False)
    l4_2 = l4_2 + l4_1
    l3_points = fp_module(r, l3_xyz, l4_xyz, l3_2, l4_2, [512, 512], "fa_layer1")
    l2_points = fp_module(r, l2_xyz, l3_xyz, l2_2, l3_points, [256, 256], "fa_layer2")
    l1_points = fp_module(r, l1_xyz, l2_xyz, l1_2, l2_points, [256, 128], "fa_layer3")
    l0_points = fp_module(r, l0_xyz, l1_xyz, l0_points, l1_points, [128, 128, 128], "fa_layer4")
    net = r.dense(l0_points, "fc1", 128, bn=True, act=lambda t: F.leaky_relu(t, 0.2))
    net = r.drop(net, 0.5)
    net = r.dense(net, "fc0", num_class, act=None)
    if label is None:
        return net, None
    ce = F.cross_entropy(net.reshape(-1, num_class), label.reshape(-1), reduction="none")
    w = smpw.reshape(-1).to(r.dtype)
    classify_loss = (ce * w).sum() / torch.count_nonzero(w).clamp(min=1)
    uniform_loss = get_repulsion_loss(r, l1_xyz, 20, radius)
    # :83 `weight_reg = tf.add_n(tf.get_collection('losses'))`: tf.losses.sparse_softmax_cross_entropy registers its
    # result in the 'losses' collection (GraphKeys.LOSSES), and no weight-decay term is ever added to it (every
    # training script leaves weight_decay=None) -> weight_reg IS classify_loss: it is counted twice (:85).
    weight_reg = classify_loss
    total = classify_loss + weight_reg + uniform_weight * uniform_loss + 1e-4 * l2_on_weights(r)
    return net, total
