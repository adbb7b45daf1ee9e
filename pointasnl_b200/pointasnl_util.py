"""Drop-in for utils/pointasnl_util.py (reference): same function names, positional order and return arity,
taking / returning torch CUDA tensors.  TF-specific trailing arguments (is_training, bn_decay, weight_decay,
scope) are kept; variables live in the active ``tf_util.VariableStore`` under the names TF would give them.

What runs where: every point-set step (FPS, kNN / ball query, neighbourhood gathers, the Adaptive-Sampling
attention + re-sampling, the Non-Local attention) is one ``pasnl_*`` CUDA kernel (include/pasnl.h); the 1x1-conv /
BN / ReLU glue between them runs on the library's tcgen05 GEMMs and fused batch-norm kernels (tf_util.py, _ops.py).
"""
import math

import torch

from . import _ops, tf_grouping, tf_interpolate, tf_sampling, tf_util
from .tf_interpolate import three_interpolate, three_nn  # noqa: F401  (re-exported like the reference module)


# ---------------------------------------------------------------------------------------------- memoisation
# FPS and the neighbour searches are pure functions of coordinates.  pointasnl_sem_seg_res runs layer1_1 and
# layer1_2 on the very same inputs (reference models/pointasnl_sem_seg_res.py:35-36: FPS 10240->1280 and kNN twice).
# The INTEGER index results (never an autograd-tracked tensor) are memoised per source tensor object, and only inside
# an explicit ``index_cache()`` scope that a caller opens around ONE forward pass (callers/pointasnl_sem_seg_res.py):
# outside a scope nothing is cached, so static input buffers, CUDA-graph capture after warm-up steps and grad-mode
# changes can never see a stale entry.
_memo = None


class index_cache:
    """``with index_cache(): model forward`` -- FPS / neighbour indices computed twice on the same tensors are reused
    within the block and dropped at its end."""

    def __enter__(self):
        global _memo
        self.prev, _memo = _memo, {}
        return self

    def __exit__(self, *exc):
        global _memo
        _memo = self.prev


def _memoized(tag, srcs, extra, fn):
    if _memo is None:
        return fn()
    key = (tag, tuple(id(t) for t in srcs), tuple(t._version for t in srcs), extra)
    hit = _memo.get(key)
    if hit is not None and all(h is t for h, t in zip(hit[0], srcs)):
        res = hit[1]
        ev = _memo.get(("pending", id(res)))
        if ev is not None:               # produced by prefetch_neighbourhood on a side stream: EVERY consumer stream joins
            torch.cuda.current_stream().wait_event(ev)
            res.record_stream(torch.cuda.current_stream())
        return res
    res = fn()
    _memo[key] = (tuple(srcs), res)     # strong refs: an id cannot be recycled while the scope is open
    return res


_side_streams = {}
_branch_streams = {}


class branch:
    """``with branch(x, y) as br: z = f(x, y)`` runs f on a side stream that first waits for the current stream (so x, y
    are ready); ``br.join(z, ...)`` makes the current stream wait for it before z is used.  Independent branches of a
    layer (the Non-Local cell next to the local cell, the two global heads of the classifier) then overlap instead of
    queueing behind each other -- most of their kernels are latency-bound and leave SMs idle.  Autograd replays each
    node's backward on the stream its forward ran on, so the backward branches overlap too.  PASNL_BRANCH_STREAMS=0
    (or CPU tensors) turns the class into a no-op."""
    enabled = None

    def __init__(self, *tensors, slot=0):
        import os
        if branch.enabled is None:
            branch.enabled = os.environ.get("PASNL_BRANCH_STREAMS", "1")[:1] != "0"
        self.active = branch.enabled and bool(tensors) and all(t.is_cuda for t in tensors if t is not None)
        self.tensors = [t for t in tensors if t is not None]
        self.slot = slot

    def __enter__(self):
        if self.active:
            dev = self.tensors[0].device
            self.main = torch.cuda.current_stream(dev)
            key = (dev.index, self.slot, self.main.cuda_stream)      # every parent stream has its own children
            self.side = _branch_streams.get(key)
            if self.side is None:
                self.side = _branch_streams[key] = torch.cuda.Stream(dev)
            self.side.wait_stream(self.main)
            for t in self.tensors:
                t.record_stream(self.side)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.active:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *outs):
        if self.active:
            self.main.wait_stream(self.side)
            for t in outs:
                if t is not None:
                    t.record_stream(self.main)


_prefetch_enabled = True


def set_concurrency(enabled):
    """Switch the side-stream prefetch and the branch streams on / off for the layers built from now on (default on;
    off = every kernel of a step on the caller's stream, e.g. to time launches one at a time)."""
    global _prefetch_enabled
    _prefetch_enabled = bool(enabled)
    import os
    branch.enabled = bool(enabled) and os.environ.get("PASNL_BRANCH_STREAMS", "1")[:1] != "0"


def prefetch_neighbourhood(xyz, npoint, nsample):
    """Launch the FPS (xyz -> npoint samples) and the k-NN (nsample neighbours of the samples in xyz) of a LATER
    set-abstraction layer NOW, on a side stream: both depend on coordinates only, FPS is a latency chain that occupies
    one CTA (or one cluster) per cloud, and the layers in between keep the other SMs busy (SemanticKITTI shape: 13.9 ms
    of FPS on 16 of 148 SMs next to ~10 ms of layer-0 work).  The results land in the open index_cache(); the layer
    that asks for them joins the side stream.  No-op outside an index_cache() scope or on CPU tensors."""
    if _memo is None or not xyz.is_cuda or not _prefetch_enabled:
        return
    dev = xyz.device
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side), torch.no_grad():
        fps_idx = _fps_idx(npoint, xyz)
        q = _ops.GatherRows.apply(xyz.detach(), fps_idx)
        idx = _neighbor_idx(nsample, xyz, q, True, None, q_key=fps_idx)
        ev = torch.cuda.Event()
        ev.record(side)
    xyz.record_stream(side)
    _memo[("pending", id(fps_idx))] = ev
    _memo[("pending", id(idx))] = ev


# ---------------------------------------------------------------------------------------------- point-set functions
def knn_query(k, support_pts, query_pts):
    """support_pts (B, N1, 3), query_pts (B, N2, 3) -> neighbor_idx (B, N2, k) int32 (reference :22-30).
    The reference leaves the device for a nanoflann kd-tree behind tf.py_func; this is one CUDA kernel with the same
    exact-kNN contract (ascending squared distance, the query itself first when it is a support point)."""
    return _ops.knn(k, support_pts, query_pts)


def _fps_idx(npoint, pts):
    return _memoized("fps", (pts,), int(npoint), lambda: tf_sampling.farthest_point_sample(npoint, pts.detach()))


def sampling(npoint, pts, feature=None):
    """pts (B, N, D) -> sub_pts (B, npoint, D) [, sub_feature (B, npoint, C)] by farthest point sampling on pts
    (reference :33-49).  Differentiable w.r.t. pts / feature (tf.gather_nd semantics)."""
    fps_idx = _fps_idx(npoint, pts)
    sub = _ops.GatherRows.apply(pts, fps_idx)
    if feature is None:
        return sub
    return sub, _ops.GatherRows.apply(feature, fps_idx)


def _neighbor_idx(K, src_xyz, q_xyz, use_knn, radius, q_key=None):
    """q_key: the tensor that identifies the query set for memoisation (the FPS index the queries were gathered with:
    two gathers of the same rows are different tensor objects but the same queries)."""
    def run():
        if use_knn:
            return knn_query(K, src_xyz.detach(), q_xyz.detach())
        return tf_grouping.query_ball_point(radius, K, src_xyz.detach(), q_xyz.detach())[0]
    return _memoized("nbr", (src_xyz, q_xyz if q_key is None else q_key), (int(K), bool(use_knn), radius), run)


def grouping(feature, K, src_xyz, q_xyz, use_xyz=True, use_knn=True, radius=0.2):
    """K neighbours of every q_xyz in src_xyz (reference :51-76) -> grouped_xyz (B, npoint, K, 3),
    grouped_feature (B, npoint, K, [3+]C), idx.

    use_knn=True is the path every shipped model takes.  use_knn=False is BROKEN in the reference (it uses `idx`
    before defining it, :67-71); here it has the obviously intended meaning (ball query + group_point).
    idx is returned as the (B, npoint, K) int32 neighbour index (the reference returns the (B, npoint, K, 2)
    [batch, point] pairs tf.gather_nd needs; no caller reads it)."""
    idx = _neighbor_idx(K, src_xyz, q_xyz, use_knn, radius)
    if use_xyz:
        grouped_feature = _ops.GroupConcat.apply(src_xyz, feature, idx, None)   # [xyz | feature] in one pass
        grouped_xyz = grouped_feature[..., :3]
    else:
        grouped_xyz = tf_grouping.group_point(src_xyz, idx)
        grouped_feature = tf_grouping.group_point(feature, idx)
    return grouped_xyz, grouped_feature, idx


# ---------------------------------------------------------------------------------------------- small MLP helpers
def weight_net_hidden(xyz, hidden_units, scope, is_training, bn_decay=None, weight_decay=None,
                      activation_fn=tf_util.relu):
    """reference :78-88"""
    with tf_util.variable_scope(scope):
        net = xyz
        for i, num_hidden_units in enumerate(hidden_units):
            net = tf_util.conv2d(net, num_hidden_units, [1, 1], padding="VALID", stride=[1, 1], bn=True,
                                 is_training=is_training, activation_fn=activation_fn, scope="wconv%d" % i,
                                 bn_decay=bn_decay, weight_decay=weight_decay)
    return net


def nonlinear_transform(data_in, mlp, scope, is_training, bn_decay=None, weight_decay=None,
                        activation_fn=tf_util.relu):
    """reference :90-110 (unused by the shipped models)"""
    with tf_util.variable_scope(scope):
        net = data_in
        n_layers = len(mlp)
        for i, out_ch in enumerate(mlp[:-1]):
            net = tf_util.conv2d(net, out_ch, [1, 1], padding="VALID", stride=[1, 1], bn=True, is_training=is_training,
                                 activation_fn=tf_util.relu, scope="nonlinear%d" % i, bn_decay=bn_decay,
                                 weight_decay=weight_decay)
        net = tf_util.conv2d(net, mlp[-1], [1, 1], padding="VALID", stride=[1, 1], bn=False, is_training=is_training,
                             scope="nonlinear%d" % (n_layers - 1), bn_decay=bn_decay, activation_fn=torch.sigmoid,
                             weight_decay=weight_decay)
    return net


# ---------------------------------------------------------------------------------------------- Adaptive Sampling
def _sample_logits(new_point, grouped_xyz, mlps, is_training, bn_decay, weight_decay, scope, bn=True, scaled=True):
    """Body of SampleWeights up to (excluding) the final softmax over the member axis."""
    with tf_util.variable_scope(scope):
        channel = new_point.shape[-1]
        bottleneck_channel = max(32, channel // 2)
        normalized_xyz = grouped_xyz - grouped_xyz[:, :, 0:1, :]
        x = torch.cat([normalized_xyz, new_point], dim=-1)          # (B, S, k, channel + 3)
        kv = tf_util.conv2d(x, bottleneck_channel * 2, [1, 1], padding="VALID", stride=[1, 1], bn=bn,
                            is_training=is_training, scope="conv_kv_ds", bn_decay=bn_decay, weight_decay=weight_decay,
                            activation_fn=None)
        q = tf_util.conv2d(x, bottleneck_channel, [1, 1], padding="VALID", stride=[1, 1], bn=bn,
                           is_training=is_training, scope="conv_query_ds", bn_decay=bn_decay,
                           weight_decay=weight_decay, activation_fn=None)
        scale = 1.0 / math.sqrt(float(bottleneck_channel)) if scaled else 1.0
        net = _ops.ASCore.apply(q, kv, scale)                        # softmax(q k^T * scale) v, fused per group
        for i, c in enumerate(mlps):
            activation = tf_util.relu if i < len(mlps) - 1 else None
            net = tf_util.conv2d(net, c, [1, 1], padding="VALID", stride=[1, 1], bn=bn, is_training=is_training,
                                 scope="mlp2_%d" % i, bn_decay=bn_decay, weight_decay=weight_decay,
                                 activation_fn=activation)
    return net


def SampleWeights(new_point, grouped_xyz, mlps, is_training, bn_decay, weight_decay, scope, bn=True, scaled=True):
    """new_point (B, S, k, C), grouped_xyz (B, S, k, 3) -> weights (B, S, k, mlps[-1]), softmax over the k members
    (reference :112-156)."""
    return torch.softmax(_sample_logits(new_point, grouped_xyz, mlps, is_training, bn_decay, weight_decay, scope, bn,
                                        scaled), dim=2)


def AdaptiveSampling(group_xyz, group_feature, num_neighbor, is_training, bn_decay, weight_decay, scope, bn):
    """group_xyz (B, S, K, 3), group_feature (B, S, K, Cg) -> new_xyz (B, S, 3), new_feature (B, S, Cg)
    (reference :158-173).  The member softmax and both weighted sums are one fused kernel."""
    with tf_util.variable_scope(scope):
        num_channel = group_feature.shape[-1]
        if num_neighbor == 0:
            return group_xyz[:, :, 0, :], group_feature[:, :, 0, :]
        shift_group_xyz = group_xyz[:, :, :num_neighbor, :]
        shift_group_points = group_feature[:, :, :num_neighbor, :]
        logits = _sample_logits(shift_group_points, shift_group_xyz, [32, 1 + num_channel], is_training, bn_decay,
                                weight_decay, scope, bn)
        new_xyz, new_feature = _ops.ASResample.apply(logits, shift_group_xyz, shift_group_points)
    return new_xyz, new_feature


# ---------------------------------------------------------------------------------------------- Non-Local cell
def PointNonLocalCell(feature, new_point, mlp, is_training, bn_decay, weight_decay, scope, bn=True, scaled=True,
                      mode="dot"):
    """feature (B, ndataset, C), new_point (B, npoint, nsample, Cq) -> (B, npoint*nsample..., mlp[-1]) exactly as the
    reference shapes it (reference :175-219; the SA layer calls it with new_point = (B, 1, npoint, Cq)).
    The attention map is never materialised: softmax(QK^T)V is one flash-style kernel."""
    if mode not in ("dot", "concat"):
        raise ValueError("mode must be 'dot' or 'concat'")
    with tf_util.variable_scope(scope):
        bottleneck_channel = mlp[0]
        batch_size, npoint, nsample, _ = new_point.shape
        kv = tf_util.conv2d(feature.unsqueeze(2), bottleneck_channel * 2, [1, 1], padding="VALID", stride=[1, 1],
                            bn=bn, is_training=is_training, scope="conv_kv", bn_decay=bn_decay,
                            weight_decay=weight_decay, activation_fn=None)
        q = tf_util.conv2d(new_point, bottleneck_channel, [1, 1], padding="VALID", stride=[1, 1], bn=bn,
                           is_training=is_training, scope="conv_query", bn_decay=bn_decay, weight_decay=weight_decay,
                           activation_fn=None)
        q = q.reshape(batch_size, npoint * nsample, bottleneck_channel)
        kv = kv.squeeze(2)
        if mode == "concat":
            out = _concat_attention(q, kv, bottleneck_channel, is_training, bn_decay, weight_decay, bn)
            out = tf_util.conv2d(out.reshape(batch_size, npoint, nsample, bottleneck_channel), mlp[-1], [1, 1],
                                 padding="VALID", stride=[1, 1], bn=bn, is_training=is_training,
                                 scope="conv_back_project", bn_decay=bn_decay, weight_decay=weight_decay)
            return out.squeeze(1)
        scale = 1.0 / math.sqrt(float(bottleneck_channel)) if scaled else 1.0
        d = bottleneck_channel
        if d % 4:   # the kernel wants a multiple of 4: zero columns change neither q.k nor the kept part of A.v
            pad = 4 - d % 4
            q = torch.nn.functional.pad(q, (0, pad))
            kv = torch.cat([torch.nn.functional.pad(kv[..., :d], (0, pad)),
                            torch.nn.functional.pad(kv[..., d:], (0, pad))], dim=-1)
        out = _ops.NLCore.apply(q, kv, scale)[..., :d]
        out = tf_util.conv2d(out.reshape(batch_size, npoint, nsample, d), mlp[-1], [1, 1], padding="VALID",
                             stride=[1, 1], bn=bn, is_training=is_training, scope="conv_back_project",
                             bn_decay=bn_decay, weight_decay=weight_decay)
        return out.squeeze(1)


def _concat_attention(q, kv, d, is_training, bn_decay, weight_decay, bn):
    """mode='concat' (reference :201-209; no shipped model uses it): attention_map[b,i,j] = relu(bn(w . [k_j | q_i] + c)),
    a 1x1 conv with ONE output channel over the tiled / concatenated (B, S, N, 2D) tensor.  The conv is linear, so the
    map is a_j + b_i with a = K w[:D], b = Q w[D:]: the 2D-wide tensor of the reference is never built; the (B, S, N)
    map itself is materialised (plain PyTorch ops: this mode is off the hot path), batch-norm statistics over all its
    B*S*N entries as tf.contrib.layers.batch_norm takes them for a one-channel NHWC tensor."""
    store = tf_util.get_default_store()
    kf, v = kv[..., :d], kv[..., d:]
    with tf_util.variable_scope("conv_attention_map"):
        w = tf_util._weights((1, 1, 2 * d, 1), 2 * d, 1, True, 1e-3, weight_decay)
        bias = store.get_variable("biases", (1,), tf_util.constant_init(0.0))
        a = torch.matmul(kf, w.view(2 * d, 1)[:d]).squeeze(-1)            # (B, N)
        b = torch.matmul(q, w.view(2 * d, 1)[d:]).squeeze(-1)             # (B, S)
        amap = (b.unsqueeze(2) + a.unsqueeze(1) + bias).unsqueeze(-1)      # (B, S, N, 1)
        if bn:
            amap = tf_util.batch_norm_for_conv2d(amap, is_training, bn_decay, "bn")
        amap = torch.relu(amap).squeeze(-1)
    return torch.matmul(torch.softmax(amap, dim=-1), v)


# ---------------------------------------------------------------------------------------------- SA layer
def PointASNLSetAbstraction(xyz, feature, npoint, nsample, mlp, is_training, bn_decay, weight_decay, scope, bn=True,
                            use_knn=True, radius=None, as_neighbor=8, NL=True, prefetch_next=None):
    """xyz (B, ndataset, 3), feature (B, ndataset, C) -> new_xyz (B, npoint, 3), new_points (B, npoint, mlp[-1])
    (reference :221-292).  Same computation, re-ordered around ONE neighbour search whose index feeds two fused
    gathers: the first-as_neighbor view for Adaptive Sampling, then [rel xyz | abs xyz | feature] for the local cell.
    prefetch_next (extra, optional): see prefetch_neighbourhood."""
    with tf_util.variable_scope(scope):
        batch_size, num_points, num_channel = feature.shape
        sampled = num_points != npoint
        if sampled:
            fps_idx = _fps_idx(npoint, xyz)                               # FPS + row gathers (sampling(), :33-49)
            new_xyz, new_feature = _ops.GatherRows.apply(xyz, fps_idx), _ops.GatherRows.apply(feature, fps_idx)
        else:
            fps_idx, new_xyz, new_feature = None, xyz, feature
        idx = _neighbor_idx(nsample, xyz, new_xyz, use_knn, radius, q_key=fps_idx)   # (B, npoint, nsample) int32
        nl_channel = mlp[-1]

        if sampled:                                                        # Adaptive Sampling (reference :246-247)
            k_as = as_neighbor if as_neighbor > 0 else 1
            group_as = _ops.GroupConcat.apply(xyz, feature, idx[:, :, :k_as], None)   # (B, S, k, 3 + C) = [xyz | feat]
            new_xyz, new_feature = AdaptiveSampling(group_as[..., :3], group_as, as_neighbor, is_training, bn_decay,
                                                    weight_decay, scope, bn)
        if prefetch_next is not None:      # (npoint, nsample) of the NEXT layer, which samples from this layer's new_xyz:
            prefetch_neighbourhood(new_xyz, *prefetch_next)   # its FPS + k-NN run beside this layer's feature work
        # translation normalisation + concat (reference :248-249): [xyz_j - new_xyz | xyz_j | feature_j]
        new_point = _ops.GroupConcat.apply(xyz, feature, idx, new_xyz)
        if nsample <= 64:
            # one autograd node for the three readers of new_point: skip-connection max-pool, conv0, weight net
            skip_spatial, new_point, grouped_xyz = _ops.SAFront.apply(new_point)
        else:
            skip_spatial = _ops.MaxPoolK.apply(new_point)
            grouped_xyz = new_point[..., :3]

        nl_branch = None
        if NL:                                                             # reference :252-256 -- on a side stream
            nl_branch = branch(feature, new_feature, slot=1)
            with nl_branch:
                new_nonlocal_point = PointNonLocalCell(feature, new_feature.unsqueeze(1),
                                                       [max(32, num_channel // 2), nl_channel], is_training, bn_decay,
                                                       weight_decay, scope, bn)
        # skip connection (reference :258-262)
        skip_branch = branch(skip_spatial, slot=4)       # a small layer (rows = B*S): beside the tall local-cell layers
        with skip_branch:
            skip_spatial = tf_util.conv1d(skip_spatial, mlp[-1], 1, padding="VALID", stride=1, bn=bn,
                                          is_training=is_training, scope="skip", bn_decay=bn_decay,
                                          weight_decay=weight_decay)
        # point local cell (reference :264-280)
        if len(mlp) == 3 and bn:
            # conv0 -> conv1 as one fused pair (conv0's activation stays a GEMM-operand prologue of conv1)
            # (folding conv1's as well, into the PointConv kernels -- _ops.BNActPointConv -- was measured and is NOT
            #  used: those kernels are shared-memory bound and lose more than the saved normalise pass)
            new_point = tf_util.conv2d_pair(new_point, mlp[:2], ("conv0", "conv1"), bn_decay=bn_decay,
                                            is_training=is_training, weight_decay=weight_decay)
        else:
            for i, num_out_channel in enumerate(mlp):
                if i != len(mlp) - 1:
                    new_point = tf_util.conv2d(new_point, num_out_channel, [1, 1], padding="VALID", stride=[1, 1], bn=bn,
                                               is_training=is_training, scope="conv%d" % i, bn_decay=bn_decay,
                                               weight_decay=weight_decay)
        weight = weight_net_hidden(grouped_xyz, [32], scope="weight_net", is_training=is_training, bn_decay=bn_decay,
                                   weight_decay=weight_decay)
        new_point = _ops.pointconv_contract(new_point, weight)             # per group (C' x K)(K x 32), one kernel
        new_point = tf_util.conv2d(new_point, mlp[-1], [1, new_point.shape[2]], padding="VALID", stride=[1, 1],
                                   bn=bn, is_training=is_training, scope="after_conv", bn_decay=bn_decay,
                                   weight_decay=weight_decay)
        new_point = new_point.squeeze(2)
        if nl_branch is not None:
            nl_branch.join(new_nonlocal_point)
        skip_branch.join(skip_spatial)
        new_point = _ops.add_n(new_point, skip_spatial, new_nonlocal_point if NL else None)
        # feature fusion (reference :287-290)
        new_point = tf_util.conv1d(new_point, mlp[-1], 1, padding="VALID", stride=1, bn=bn, is_training=is_training,
                                   scope="aggregation", bn_decay=bn_decay, weight_decay=weight_decay)
        return new_xyz, new_point


# ---------------------------------------------------------------------------------------------- decoding layer
def PointASNLDecodingLayer(xyz1, xyz2, points1, points2, nsample, mlp, is_training, bn_decay, weight_decay, scope,
                           bn=True, use_xyz=True, use_knn=True, radius=None, dilate_rate=1, mode="concat", NL=False):
    """xyz1 (B, n1, 3) dense, xyz2 (B, n2, 3) sparse, points1 (B, n1, c1), points2 (B, n2, c2) -> (B, n1, mlp[-1])
    (reference :294-351; used by models/pointasnl_sem_seg.py)."""
    with tf_util.variable_scope(scope):
        num_channel = points2.shape[-1]
        dist, idx = three_nn(xyz1.detach(), xyz2.detach())
        dist = torch.clamp(dist, min=1e-10)
        norm = (1.0 / dist).sum(dim=2, keepdim=True)
        weight = (1.0 / dist) / norm
        if NL:
            new_nonlocal_point = PointNonLocalCell(points1, points2.unsqueeze(1), [max(32, num_channel), num_channel],
                                                   is_training, bn_decay, weight_decay, scope, bn, mode=mode)
            points2 = _ops.add_n(points2, new_nonlocal_point)
        interpolated_points = three_interpolate(points2, idx, weight)
        grouped_xyz, grouped_feature, _ = grouping(interpolated_points, nsample, xyz1, xyz1, use_xyz=use_xyz,
                                                   use_knn=use_knn, radius=radius)
        grouped_xyz = grouped_xyz - xyz1.unsqueeze(2)
        weight = weight_net_hidden(grouped_xyz, [32], scope="decode_weight_net", is_training=is_training,
                                   bn_decay=bn_decay, weight_decay=weight_decay)
        new_points = _ops.pointconv_contract(grouped_feature, weight)
        new_points = tf_util.conv2d(new_points, mlp[0], [1, new_points.shape[2]], padding="VALID", stride=[1, 1],
                                    bn=bn, is_training=is_training, scope="decode_after_conv", bn_decay=bn_decay,
                                    weight_decay=weight_decay)
        if points1 is not None:
            new_points1 = _ops.concat_cols(new_points, points1.unsqueeze(2))
        else:
            new_points1 = new_points
        for i, num_out_channel in enumerate(mlp):
            if i != 0:
                new_points1 = tf_util.conv2d(new_points1, num_out_channel, [1, 1], padding="VALID", stride=[1, 1],
                                             bn=bn, is_training=is_training, scope="conv_%d" % i, bn_decay=bn_decay,
                                             weight_decay=weight_decay)
        return new_points1.squeeze(2)


# ---------------------------------------------------------------------------------------------- repulsion loss
def get_repulsion_loss(pred, nsample=20, radius=0.07):
    """pred (B, npoint, 3) -> scalar (reference :361-378): ball query + group on (pred, pred), the 4 smallest
    non-self squared distances per point, mean(radius - d * exp(-d^2 / h^2)), h = 0.03."""
    idx, _pts_cnt = tf_grouping.query_ball_point(radius, nsample, pred.detach(), pred.detach())
    grouped_pred = tf_grouping.group_point(pred, idx)
    grouped_pred = grouped_pred - pred.unsqueeze(2)
    h = 0.03
    dist_square = (grouped_pred ** 2).sum(dim=-1)
    neg, _ = torch.topk(-dist_square, 5, dim=-1)
    dist_square = -neg[:, :, 1:]
    dist_square = torch.clamp(dist_square, min=1e-12)
    dist = torch.sqrt(dist_square)
    weight = torch.exp(-dist_square / h ** 2)
    return torch.mean(radius - dist * weight)
