"""torch-facing wrappers of the C ABI: argument validation, output allocation, autograd.

PyTorch is used here for device memory, streams and autograd plumbing only; every computation below is one
``pasnl_*`` call into libpasnl_b200.so (include/pasnl.h).  No hot-path op (sections "sampling" ... "Non-Local core")
has a CPU or pure-PyTorch fallback: a non-CUDA tensor is an error.

Gradient contract = the reference's registrations: gather_point -> inp (tf_sampling.py:43-47),
group_point -> points (tf_grouping.py:42-46), three_interpolate -> points only (tf_interpolate.py:29-34); no
gradient for FPS / ball query / kNN / three_nn (ops.NoGradient at tf_sampling.py:22,57; tf_grouping.py:21,32;
tf_interpolate.py:18).
"""
import torch

from . import _lib

_call = _lib.call
_ptr = _lib.ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_current_device(t, name):
    """Kernels are launched on the CURRENT device's current stream with raw pointers: a tensor living on another
    device would be dereferenced on the wrong GPU (illegal address at best).  Fail loudly instead."""
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise ValueError(f"{name} is on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "wrap the call in `with torch.cuda.device(tensor.device):`")


def _f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: pointasnl_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    _on_current_device(t, name)
    return t.contiguous()


def _i32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: pointasnl_b200 has no CPU path (got device {t.device})")
    if t.dtype != torch.int32:
        if t.dtype in (torch.int64, torch.int16, torch.uint8):
            t = t.to(torch.int32)
        else:
            raise ValueError(f"{name} must be an integer tensor, got {t.dtype}")
    _on_current_device(t, name)
    return t.contiguous()


def _require(cond, msg):
    if not cond:
        raise ValueError(msg)  # mirrors TF's errors::InvalidArgument at the op boundary


# ------------------------------------------------------------------ sampling
def fps(npoint, inp):
    inp = _f32(inp, "inp")
    _require(int(npoint) > 0, "FarthestPointSample expects positive npoint")
    _require(inp.dim() == 3 and inp.shape[2] == 3, "FarthestPointSample expects (batch_size,num_points,3) inp shape")
    b, n, _ = inp.shape
    _require(n >= 1, "FarthestPointSample expects at least one point per cloud")
    out = torch.empty((b, int(npoint)), device=inp.device, dtype=torch.int32)
    ws_bytes = _lib.load().pasnl_fps_workspace_bytes(b, n, int(npoint))
    ws = torch.empty((ws_bytes,), device=inp.device, dtype=torch.uint8) if ws_bytes else None
    _call("pasnl_fps", b, n, int(npoint), _ptr(inp), _ptr(ws), _ptr(out), _stream())
    return out


class GatherPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, idx):
        inp, idx = _f32(inp, "inp"), _i32(idx, "idx")
        _require(inp.dim() == 3 and inp.shape[2] == 3, "GatherPoint expects (batch_size,num_points,3) inp shape")
        _require(idx.dim() == 2 and idx.shape[0] == inp.shape[0], "GatherPoint expects (batch_size,num_result) idx shape")
        b, n, _ = inp.shape
        m = idx.shape[1]
        out = torch.empty((b, m, 3), device=inp.device, dtype=torch.float32)
        _call("pasnl_gather_point", b, n, m, _ptr(inp), _ptr(idx), _ptr(out), _stream())
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, out_g):
        (idx,) = ctx.saved_tensors
        out_g = _f32(out_g, "out_g")
        b, m = idx.shape
        inp_g = torch.empty((b, ctx.n, 3), device=out_g.device, dtype=torch.float32)
        _call("pasnl_gather_point_grad", b, ctx.n, m, _ptr(out_g), _ptr(idx), _ptr(inp_g), _stream())
        return inp_g, None


class GatherRows(torch.autograd.Function):
    """out[b,j,:] = src[b,idx[b,j],:] for any channel count (tf.gather_nd of sampling(), pointasnl_util.py:43-49)."""

    @staticmethod
    def forward(ctx, src, idx):
        src, idx = _f32(src, "src"), _i32(idx, "idx")
        _require(src.dim() == 3, "gather expects (batch_size,num_points,channel) src shape")
        _require(idx.dim() == 2 and idx.shape[0] == src.shape[0], "gather expects (batch_size,num_result) idx shape")
        b, n, c = src.shape
        m = idx.shape[1]
        out = torch.empty((b, m, c), device=src.device, dtype=torch.float32)
        _call("pasnl_gather_rows", b, n, c, m, _ptr(src), _ptr(idx), _ptr(out), _stream())
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, out_g):
        (idx,) = ctx.saved_tensors
        out_g = _f32(out_g, "out_g")
        b, m, c = out_g.shape
        src_g = torch.empty((b, ctx.n, c), device=out_g.device, dtype=torch.float32)
        _call("pasnl_gather_rows_grad", b, ctx.n, c, m, _ptr(out_g), _ptr(idx), _ptr(src_g), _stream())
        return src_g, None


def prob_sample(inp, inpr):
    inp, inpr = _f32(inp, "inp"), _f32(inpr, "inpr")
    _require(inp.dim() == 2, "ProbSample expects (batch_size,num_choices) inp shape")
    _require(inpr.dim() == 2 and inpr.shape[0] == inp.shape[0], "ProbSample expects (batch_size,num_points) inpr shape")
    b, n = inp.shape
    m = inpr.shape[1]
    temp = torch.empty((b, n), device=inp.device, dtype=torch.float32)
    out = torch.empty((b, m), device=inp.device, dtype=torch.int32)
    _call("pasnl_prob_sample", b, n, m, _ptr(inp), _ptr(inpr), _ptr(temp), _ptr(out), _stream())
    return out


def selection_sort(k, dist):
    dist = _f32(dist, "dist")
    _require(dist.dim() == 3, "SelectionSort expects (b,m,n) dist shape.")
    _require(int(k) > 0, "SelectionSort expects positive k")
    b, m, n = dist.shape
    outi = torch.empty((b, m, n), device=dist.device, dtype=torch.int32)
    out = torch.empty((b, m, n), device=dist.device, dtype=torch.float32)
    _call("pasnl_selection_sort", b, n, m, int(k), _ptr(dist), _ptr(outi), _ptr(out), _stream())
    return outi, out


# ------------------------------------------------------------------ grouping
def ball_query(radius, nsample, xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1, "xyz1"), _f32(xyz2, "xyz2")
    _require(float(radius) > 0, "QueryBallPoint expects positive radius")
    _require(int(nsample) > 0, "QueryBallPoint expects positive nsample")
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0],
             "QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.empty((b, m, int(nsample)), device=xyz1.device, dtype=torch.int32)
    cnt = torch.empty((b, m), device=xyz1.device, dtype=torch.int32)
    _call("pasnl_ball_query", b, n, m, float(radius), int(nsample), _ptr(xyz1), _ptr(xyz2), _ptr(idx), _ptr(cnt),
          _stream())
    return idx, cnt


class GroupPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        points, idx = _f32(points, "points"), _i32(idx, "idx")
        _require(points.dim() == 3, "GroupPoint expects (batch_size, num_points, channel) points shape")
        _require(idx.dim() == 3 and idx.shape[0] == points.shape[0],
                 "GroupPoint expects (batch_size, npoints, nsample) idx shape")
        b, n, c = points.shape
        _, m, ns = idx.shape
        out = torch.empty((b, m, ns, c), device=points.device, dtype=torch.float32)
        _call("pasnl_group_point", b, n, c, m, ns, _ptr(points), _ptr(idx), _ptr(out), _stream())
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = _f32(grad_out, "grad_out")
        b, m, ns, c = grad_out.shape
        g = torch.empty((b, ctx.n, c), device=grad_out.device, dtype=torch.float32)
        _call("pasnl_group_point_grad", b, ctx.n, c, m, ns, _ptr(grad_out), _ptr(idx), _ptr(g), _stream())
        return g, None


def knn(k, xyz1, xyz2, return_dist=False):
    """k nearest support points (xyz1) of every query (xyz2), ascending (squared distance, index)."""
    xyz1, xyz2 = _f32(xyz1, "xyz1"), _f32(xyz2, "xyz2")
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "knn expects (batch_size, ndataset, 3) support shape")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0],
             "knn expects (batch_size, npoint, 3) query shape")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    k = int(k)
    _require(1 <= k <= n, f"knn expects 1 <= k <= ndataset (k={k}, ndataset={n})")
    _require(k <= 128, "knn supports k <= 128")
    idx = torch.empty((b, m, k), device=xyz1.device, dtype=torch.int32)
    dist = torch.empty((b, m, k), device=xyz1.device, dtype=torch.float32) if return_dist else None
    _call("pasnl_knn", b, n, m, k, _ptr(xyz1), _ptr(xyz2), _ptr(idx), _ptr(dist), _stream())
    return (dist, idx) if return_dist else idx


class GroupConcat(torch.autograd.Function):
    """[xyz[idx]-center | xyz[idx] | feat[idx]] in one pass (grouping() + translation normalisation + concat,
    pointasnl_util.py:63-74,248-249).  ``idx`` may be a last-dim slice ``full_idx[:, :, :k]`` (no copy is made)."""

    @staticmethod
    def forward(ctx, xyz, feat, idx, center):
        xyz = _f32(xyz, "xyz")
        feat = _f32(feat, "feat") if feat is not None else None
        center = _f32(center, "center") if center is not None else None
        _require(idx.is_cuda and idx.dtype == torch.int32 and idx.dim() == 3, "idx must be a CUDA int32 (b,m,nsample) tensor")
        _require(idx.stride(2) == 1 and idx.stride(0) == idx.shape[1] * idx.stride(1),
                 "idx must be contiguous or a last-dim prefix slice of a contiguous tensor")
        b, n, _ = xyz.shape
        _, m, ns = idx.shape
        c = 0 if feat is None else feat.shape[2]
        if feat is not None:
            _require(feat.shape[0] == b and feat.shape[1] == n, "feat must be (batch_size, ndataset, channel)")
        if center is not None:
            _require(tuple(center.shape) == (b, m, 3), "center must be (batch_size, npoint, 3)")
        w = (3 if center is not None else 0) + 3 + c
        out = torch.empty((b, m, ns, w), device=xyz.device, dtype=torch.float32)
        _call("pasnl_group_concat", b, n, c, m, ns, _ptr(xyz), _ptr(feat), idx.data_ptr(), int(idx.stride(1)),
              _ptr(center), _ptr(out), _stream())
        ctx.save_for_backward(idx)
        ctx.dims = (b, n, c, m, ns, center is not None, feat is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        b, n, c, m, ns, has_center, has_feat = ctx.dims
        grad_out = _f32(grad_out, "grad_out")
        dev = grad_out.device
        g_xyz = torch.empty((b, n, 3), device=dev, dtype=torch.float32)
        g_feat = torch.empty((b, n, c), device=dev, dtype=torch.float32) if has_feat else None
        g_center = torch.empty((b, m, 3), device=dev, dtype=torch.float32) if has_center else None
        _call("pasnl_group_concat_grad", b, n, c, m, ns, _ptr(grad_out), idx.data_ptr(), int(idx.stride(1)),
              1 if has_center else 0, _ptr(g_xyz), _ptr(g_feat), _ptr(g_center), _stream())
        return g_xyz, g_feat, None, g_center


class PointConvContract(torch.autograd.Function):
    """x (..., k, c), w (..., k, 32) -> x^T w per neighbourhood, (..., c, 32)  (pointasnl_util.py:272-273)."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = _f32(x, "x"), _f32(w, "w")
        k, c = x.shape[-2], x.shape[-1]
        _require(w.shape[-1] == 32 and w.shape[:-1] == x.shape[:-1], "weight must be (..., k, 32) matching x (..., k, c)")
        groups = x.numel() // (k * c)
        out = torch.empty(x.shape[:-2] + (c, 32), device=x.device, dtype=torch.float32)
        _call("pasnl_pointconv_fwd", groups, k, c, _ptr(x), _ptr(w), _ptr(out), _stream())
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = _f32(g, "grad_out")
        k, c = x.shape[-2], x.shape[-1]
        gx, gw = torch.empty_like(x), torch.empty_like(w)
        _call("pasnl_pointconv_bwd", x.numel() // (k * c), k, c, _ptr(x), _ptr(w), _ptr(g), _ptr(gx), _ptr(gw), _stream())
        return gx, gw


def pointconv_contract(x, w):
    k, c = x.shape[-2], x.shape[-1]
    if x.is_cuda and w.shape[-1] == 32 and pointconv_eligible(k, c):
        return PointConvContract.apply(x, w)
    return torch.matmul(x.transpose(-1, -2), w)


class MaxPoolK(torch.autograd.Function):
    """x (..., k, c) -> max over the k axis (..., c); gradient split evenly among tied maxima (tf.reduce_max)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "x")
        k, c = x.shape[-2], x.shape[-1]
        groups = x.numel() // (k * c)
        out = torch.empty(x.shape[:-2] + (c,), device=x.device, dtype=torch.float32)
        _call("pasnl_maxpool_k", groups, k, c, _ptr(x), _ptr(out), _stream())
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        g = _f32(g, "grad_out")
        k, c = x.shape[-2], x.shape[-1]
        gx = torch.empty_like(x)
        _call("pasnl_maxpool_k_grad", x.numel() // (k * c), k, c, _ptr(x), _ptr(out), _ptr(g), _ptr(gx), _stream())
        return gx


# ------------------------------------------------------------------ interpolation
def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1, "xyz1"), _f32(xyz2, "xyz2")
    _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "ThreeNN expects (b,n,3) xyz1 shape.")
    _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0], "ThreeNN expects (b,m,3) xyz2 shape.")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = torch.empty((b, n, 3), device=xyz1.device, dtype=torch.float32)
    idx = torch.empty((b, n, 3), device=xyz1.device, dtype=torch.int32)
    _call("pasnl_three_nn", b, n, m, _ptr(xyz1), _ptr(xyz2), _ptr(dist), _ptr(idx), _stream())
    return dist, idx


class ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, weight):
        points, idx, weight = _f32(points, "points"), _i32(idx, "idx"), _f32(weight, "weight")
        _require(points.dim() == 3, "ThreeInterpolate expects (b,m,c) points shape")
        b, m, c = points.shape
        _require(idx.dim() == 3 and idx.shape[0] == b and idx.shape[2] == 3, "ThreeInterpolate expects (b,n,3) idx shape")
        n = idx.shape[1]
        _require(tuple(weight.shape) == (b, n, 3), "ThreeInterpolate expects (b,n,3) weight shape")
        out = torch.empty((b, n, c), device=points.device, dtype=torch.float32)
        _call("pasnl_three_interpolate", b, m, c, n, _ptr(points), _ptr(idx), _ptr(weight), _ptr(out), _stream())
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        grad_out = _f32(grad_out, "grad_out")
        b, n, c = grad_out.shape
        g = torch.empty((b, ctx.m, c), device=grad_out.device, dtype=torch.float32)
        _call("pasnl_three_interpolate_grad", b, n, c, ctx.m, _ptr(grad_out), _ptr(idx), _ptr(weight), _ptr(g), _stream())
        return g, None, None  # the reference registers a gradient w.r.t. points only (tf_interpolate.py:29-34)


# ------------------------------------------------------------------ Adaptive Sampling cores
class ASCore(torch.autograd.Function):
    """softmax(Q Kf^T * scale) V per neighbourhood (pointasnl_util.py:137-146).  q (..., k, d), kv (..., k, 2d)."""

    @staticmethod
    def forward(ctx, q, kv, scale):
        q, kv = _f32(q, "q"), _f32(kv, "kv")
        k, d = q.shape[-2], q.shape[-1]
        _require(kv.shape[-2] == k and kv.shape[-1] == 2 * d and kv.shape[:-2] == q.shape[:-2], "kv must be (..., k, 2d)")
        groups = q.numel() // (k * d)
        out = torch.empty_like(q)
        attn = torch.empty(q.shape[:-1] + (k,), device=q.device, dtype=torch.float32)
        _call("pasnl_as_core_fwd", groups, k, d, _ptr(q), _ptr(kv), float(scale), _ptr(out), _ptr(attn), _stream())
        ctx.save_for_backward(q, kv, attn)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        q, kv, attn = ctx.saved_tensors
        grad_out = _f32(grad_out, "grad_out")
        k, d = q.shape[-2], q.shape[-1]
        groups = q.numel() // (k * d)
        gq, gkv = torch.empty_like(q), torch.empty_like(kv)
        _call("pasnl_as_core_bwd", groups, k, d, _ptr(q), _ptr(kv), _ptr(attn), ctx.scale, _ptr(grad_out), _ptr(gq),
              _ptr(gkv), _stream())
        return gq, gkv, None


def _member_strides(t):
    """(group_stride, member_stride) in floats of a (b, s, k, c) view whose channel dim is unit-stride and whose
    (b, s) dims collapse to one group index."""
    _require(t.dim() == 4 and t.stride(3) == 1 and t.stride(0) == t.shape[1] * t.stride(1),
             "grouped tensor must be (b, s, k, c) with unit channel stride and collapsible (b, s)")
    return int(t.stride(1)), int(t.stride(2))


class ASResample(torch.autograd.Function):
    """softmax over the k members + weighted re-sampling (pointasnl_util.py:155,168-171).
    logits (b,s,k,1+cg); xyz (b,s,k,3) and feat (b,s,k,cg) may be strided views of wider grouped tensors."""

    @staticmethod
    def forward(ctx, logits, xyz, feat):
        logits = _f32(logits, "logits")
        _require(xyz.is_cuda and feat.is_cuda and xyz.dtype == torch.float32 and feat.dtype == torch.float32,
                 "xyz / feat must be CUDA float32")
        b, s, k, cw = logits.shape
        cg = cw - 1
        _require(tuple(xyz.shape) == (b, s, k, 3) and tuple(feat.shape) == (b, s, k, cg), "xyz / feat shape mismatch")
        xgs, xms = _member_strides(xyz)
        fgs, fms = _member_strides(feat)
        groups = b * s
        new_xyz = torch.empty((b, s, 3), device=logits.device, dtype=torch.float32)
        new_feat = torch.empty((b, s, cg), device=logits.device, dtype=torch.float32)
        w = torch.empty_like(logits)
        _call("pasnl_as_resample_fwd", groups, k, cg, _ptr(logits), xyz.data_ptr(), xgs, xms, feat.data_ptr(), fgs, fms,
              _ptr(new_xyz), _ptr(new_feat), _ptr(w), _stream())
        ctx.save_for_backward(w, xyz, feat)
        return new_xyz, new_feat

    @staticmethod
    def backward(ctx, g_xyz, g_feat):
        w, xyz, feat = ctx.saved_tensors
        b, s, k, cw = w.shape
        cg = cw - 1
        g_xyz, g_feat = _f32(g_xyz, "g_xyz"), _f32(g_feat, "g_feat")
        xgs, xms = _member_strides(xyz)
        fgs, fms = _member_strides(feat)
        gl = torch.empty_like(w)
        gx = torch.empty((b, s, k, 3), device=w.device, dtype=torch.float32)
        gf = torch.empty((b, s, k, cg), device=w.device, dtype=torch.float32)
        _call("pasnl_as_resample_bwd", b * s, k, cg, _ptr(w), xyz.data_ptr(), xgs, xms, feat.data_ptr(), fgs, fms,
              _ptr(g_xyz), _ptr(g_feat), _ptr(gl), _ptr(gx), _ptr(gf), _stream())
        return gl, gx, gf


# ------------------------------------------------------------------ Non-Local core
class NLCore(torch.autograd.Function):
    """softmax(Q Kf^T * scale) V over all support points without materialising the attention map
    (pointasnl_util.py:195-212).  q (b,s,d), kv (b,n,2d) -> (b,s,d)."""

    @staticmethod
    def forward(ctx, q, kv, scale):
        q, kv = _f32(q, "q"), _f32(kv, "kv")
        _require(q.dim() == 3 and kv.dim() == 3 and kv.shape[0] == q.shape[0] and kv.shape[2] == 2 * q.shape[2],
                 "NL core expects q (b,s,d) and kv (b,n,2d)")
        b, s, d = q.shape
        n = kv.shape[1]
        _require(d % 4 == 0 and 4 <= d <= 128, "NL core supports bottleneck widths d % 4 == 0, 4 <= d <= 128")
        out = torch.empty_like(q)
        lse = torch.empty((b, s), device=q.device, dtype=torch.float32)
        _call("pasnl_nl_core_fwd", b, s, n, d, _ptr(q), _ptr(kv), float(scale), _ptr(out), _ptr(lse), _stream())
        ctx.save_for_backward(q, kv, out, lse)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        q, kv, out, lse = ctx.saved_tensors
        grad_out = _f32(grad_out, "grad_out")
        b, s, d = q.shape
        n = kv.shape[1]
        gq, gkv = torch.empty_like(q), torch.empty_like(kv)
        ws_bytes = _lib.load().pasnl_nl_core_bwd_workspace_bytes(b, s, n, d)
        ws = torch.empty((max(ws_bytes, 4),), device=q.device, dtype=torch.uint8)
        _call("pasnl_nl_core_bwd", b, s, n, d, _ptr(q), _ptr(kv), _ptr(out), _ptr(lse), ctx.scale, _ptr(grad_out),
              _ptr(gq), _ptr(gkv), _ptr(ws), _stream())
        return gq, gkv, None


def gemm_wgrad(x, dy):
    """x (rows, m), dy (rows, n) -> x^T dy (m, n): the weight gradient.  rows <= 64 (fully connected layers: rows = batch
    size): pasnl_wgrad_skinny; otherwise the tcgen05 kernels behind pasnl_gemm_wgrad_tf32x3 (TMEM-fed / short /
    narrow-operand variants, chosen by the library)."""
    rows, m = x.shape
    n = dy.shape[1]
    if x.stride(1) != 1:
        x = x.contiguous()
    if dy.stride(1) != 1:
        dy = dy.contiguous()
    if rows <= 64:
        dw = torch.empty((m, n), device=x.device, dtype=torch.float32)
        _call("pasnl_wgrad_skinny", rows, m, n, _ptr(x), int(x.stride(0)), _ptr(dy), int(dy.stride(0)), _ptr(dw), n, _stream())
        return dw
    dw = _zero_arena.zeros((m, n), torch.float32) if (_zero_arena is not None and _zero_arena.device == x.device) else None
    zeroed = dw is not None
    if dw is None:
        dw = torch.empty((m, n), device=x.device, dtype=torch.float32)
    _call("pasnl_gemm_wgrad_tf32x3", rows, m, n, _ptr(x), int(x.stride(0)), _ptr(dy), int(dy.stride(0)), _ptr(dw), n,
          1 if zeroed else 0, _stream())
    if zeroed:
        _lib.launch_count -= 1          # no memset node was enqueued
    return dw


# ------------------------------------------------------------------ dense glue: GEMM + fused BN/activation epilogue
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


class ZeroArena:
    """One device buffer that is cleared ONCE per training step and handed out in slices wherever the dense glue needs
    zero-initialised memory (batch-norm statistic accumulators, weight-gradient outputs that are reduced into with
    atomics, the identically-zero bias gradient).  A 26-layer step otherwise enqueues ~100 separate two-microsecond
    fill / memset nodes, each one a dependent launch on the critical path.

    Use:  arena = ZeroArena(device);  at the start of every step  `arena.reset()`;  run the step inside
    `with use_zero_arena(arena):`.  A request the arena cannot serve (first steps, while it is still being sized) falls
    back to a fresh torch.zeros; `reset()` then grows the buffer to the high-water mark.  Slices stay valid until the
    next reset(), i.e. gradients must be consumed (copied into the gradient bucket, applied) within the step."""

    ALIGN = 256

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = torch.zeros((0,), dtype=torch.uint8, device=self.device)
        self.off = 0
        self.need = 0

    def reset(self):
        if self.need > self.buf.numel():
            self.buf = torch.zeros((self.need,), dtype=torch.uint8, device=self.device)
        else:
            self.buf.zero_()
        self.off = 0
        self.need = 0

    def zeros(self, shape, dtype):
        numel = 1
        for d in shape:
            numel *= int(d)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        span = (nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.need += span
        if self.off + span > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + nbytes].view(dtype).view(shape)
        self.off += span
        return t


_zero_arena = None


class use_zero_arena:
    def __init__(self, arena):
        self.arena = arena

    def __enter__(self):
        global _zero_arena
        self.prev, _zero_arena = _zero_arena, self.arena
        return self.arena

    def __exit__(self, *exc):
        global _zero_arena
        _zero_arena = self.prev


def _zeros(shape, dtype, device):
    """(tensor, from_arena): zero-initialised memory, from the active arena when there is one."""
    if _zero_arena is not None and _zero_arena.device == torch.device(device):
        t = _zero_arena.zeros(tuple(shape), dtype)
        if t is not None:
            return t, True
    return torch.zeros(tuple(shape), dtype=dtype, device=device), False

import os as _os

GEMM_K_SLAB = 1024


class WeightTransposer:
    """K-major ([n][k]) copies of the layers' [k][n] weight matrices for the tcgen05 kernels that want them, refreshed
    for ALL registered weights by ONE pasnl_transpose_batched launch per optimizer step (eager PyTorch: one
    `w.t().contiguous()` kernel per layer per step).  Weights that are not registered are transposed on demand by a
    single-matrix launch of the same kernel."""

    def __init__(self):
        self.by_ptr = {}        # data_ptr of the [k][n] weight -> (k, n, transposed tensor, fresh flag holder)
        self.table = None
        self.count = 0
        self.max_tiles = 1
        self.managed = False

    def register(self, weights):
        """weights: list of parameters viewable as (k, n) = (prod(shape[:-1]), shape[-1]); call after the parameters
        have their final storage (TrainStep: after flatten_parameters)."""
        weights = list(weights)
        rows = []
        for w in weights:
            n = w.shape[-1]
            k = w.numel() // n
            wt = torch.empty((n, k), device=w.device, dtype=torch.float32)
            self.by_ptr[w.data_ptr()] = (k, n, wt)
            rows.append([w.data_ptr(), wt.data_ptr(), k, n])
            self.max_tiles = max(self.max_tiles, ((k + 31) // 32) * ((n + 31) // 32))
        if rows:
            self.table = torch.tensor(rows, dtype=torch.int64, device=weights[0].device)
            self.count = len(rows)
            self.managed = True

    def refresh(self):
        if self.count:
            _call("pasnl_transpose_batched", self.count, _ptr(self.table), int(self.max_tiles), _stream())

    def lookup(self, w2d):
        hit = self.by_ptr.get(w2d.data_ptr())
        if hit is not None and hit[0] == w2d.shape[0] and hit[1] == w2d.shape[1]:
            return hit[2]
        return None


_transposer = None
_single_tables = {}


def set_weight_transposer(t):
    global _transposer
    _transposer = t


def transposed_weight(w2d):
    """[n][k] copy of a [k][n] weight: the step's batched refresh if the weight is registered, else one launch now."""
    if _transposer is not None:
        wt = _transposer.lookup(w2d)
        if wt is not None:
            return wt
    k, n = w2d.shape
    if not w2d.is_contiguous():
        w2d = w2d.contiguous()
    wt = torch.empty((n, k), device=w2d.device, dtype=torch.float32)
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("weight used inside CUDA-graph capture without a registered WeightTransposer entry")
    table = torch.tensor([[w2d.data_ptr(), wt.data_ptr(), k, n]], dtype=torch.int64, device=w2d.device)
    _call("pasnl_transpose_batched", 1, _ptr(table), ((k + 31) // 32) * ((n + 31) // 32), _stream())
    return wt


def gemm_rows(a, w, w_kn=False, col_stats=None):
    """a (rows, k) times a weight -> (rows, n), fp32-accurate (3xTF32 on tcgen05, fp32 FMA for rows <= 64).
    w is [n][k] (w_kn=False: the transposed form, e.g. the stored weight seen from the input-gradient product) or [k][n]
    (w_kn=True: the layer's stored weight).  Routing: rows <= 64 -> pasnl_gemm_skinny; tall layers with k, n <= 128, and
    k <= 128 with many output columns (bound by writing C) -> pasnl_gemm_rows_ws (TMA-fed, weights stationary in TMEM;
    takes either layout); everything else ->
    pasnl_gemm_rows_tf32x3, which wants [n][k] (a [k][n] weight goes through the step's batched transpose).
    col_stats: optional zeroed (2n,) float64 tensor; when the kernel's epilogue can accumulate the per-column sum / sum
    of squares of the result into it the function returns (c, True), else (c, False)."""
    rows, k = a.shape
    n = w.shape[1] if w_kn else w.shape[0]
    if a.stride(1) != 1:
        a = a.contiguous()
    if w.stride(1) != 1:
        w = w.contiguous()
    if rows <= 64:
        c, _ = _zeros((rows, n), torch.float32, a.device)
        _call("pasnl_gemm_skinny", rows, k, n, _ptr(a), int(a.stride(0)), _ptr(w), int(w.stride(0)), 0 if w_kn else 1,
              _ptr(c), n, _stream())
        return c if col_stats is None else (c, False)
    lib = _lib.load()
    if (lib.pasnl_gemm_rows_ws_eligible(rows, k, n, a.data_ptr(), int(a.stride(0)))
            or lib.pasnl_gemm_rows_ws_wide_eligible(rows, k, n, a.data_ptr(), int(a.stride(0)))):
        c = torch.empty((rows, n), device=a.device, dtype=torch.float32)
        _call("pasnl_gemm_rows_ws", rows, k, n, _ptr(a), int(a.stride(0)), _ptr(w), int(w.stride(0)), 1 if w_kn else 0,
              _ptr(c), n, _ptr(col_stats), None, None, 0, _stream())
        return c if col_stats is None else (c, True)
    if (col_stats is None and not w_kn and 128 < n <= 160
            and lib.pasnl_gemm_rows_ws_eligible(rows, k, 128, a.data_ptr(), int(a.stride(0)))):
        # a tall input-gradient product with a few more than 128 output columns (the 134-wide grouped tensor of cls
        # layer 2): 128 columns through the weights-stationary kernel, the remainder through the one-tile kernel, both
        # writing their column range of the same C
        c = torch.empty((rows, n), device=a.device, dtype=torch.float32)
        _call("pasnl_gemm_rows_ws", rows, k, 128, _ptr(a), int(a.stride(0)), _ptr(w), int(w.stride(0)), 0, _ptr(c), n,
              None, None, None, 0, _stream())
        rest = w[128:]
        _call("pasnl_gemm_rows_tf32x3", rows, k, n - 128, _ptr(a), int(a.stride(0)), rest.data_ptr(), int(w.stride(0)),
              c.data_ptr() + 4 * 128, n, None, _stream())
        return c
    bt = transposed_weight(w) if w_kn else w
    # (contractions longer than 1024 run inside the library as side-by-side <= 1024-wide slabs added into a zeroed C:
    #  the TMEM accumulator truncates on every add, ~3e-8 * k relative; statistics are then left to the BN kernel)
    fuse_stats = col_stats is not None and (k <= GEMM_K_SLAB or bool(lib.pasnl_gemm_rows_tma_eligible(
        rows, k, n, a.data_ptr(), int(a.stride(0)), bt.data_ptr(), int(bt.stride(0)))))
    c = torch.empty((rows, n), device=a.device, dtype=torch.float32)
    _call("pasnl_gemm_rows_tf32x3", rows, k, n, _ptr(a), int(a.stride(0)), _ptr(bt), int(bt.stride(0)), _ptr(c), n,
          _ptr(col_stats) if fuse_stats else None, _stream())
    return c if col_stats is None else (c, fuse_stats)


def _bn_single_launch(rows, c):
    """Mirror of dense_glue.cu's choice: small layers run clear + reduce + apply as ONE cooperative launch."""
    if _os.environ.get("PASNL_BN_COOP", "1")[:1] == "0":
        return False
    lanes = c // 4 if c % 4 == 0 else c
    return rows * c <= (4 << 20) and lanes <= 256


def _rows_view(x, cin):
    """(rows, cin) view of x with a unit channel stride and ONE row pitch (a leading-column slice such as
    grouped_xyz = new_point[..., :3] qualifies: the GEMM kernels take a row pitch); a contiguous copy otherwise."""
    if x.dim() >= 2 and x.stride(-1) == 1 and x.shape[-1] == cin:
        dims = [(x.shape[d], x.stride(d)) for d in range(x.dim() - 1) if x.shape[d] != 1]
        ok = all(dims[i][1] == dims[i + 1][1] * dims[i + 1][0] for i in range(len(dims) - 1))
        if ok:
            pitch = dims[-1][1] if dims else cin
            if pitch >= cin:
                return x.as_strided((x.numel() // cin, cin), (pitch, 1))
    return x.reshape(-1, cin).contiguous()


class DenseBNAct(torch.autograd.Function):
    """act(batch_norm_train(x @ w + bias)) for a 1x1 conv / fully connected layer (tf_util.py:120-185,512-531).
    Forward: y = x.W (gemm_rows: tcgen05 3xTF32 for the tall layers, whose epilogue also accumulates the batch-norm
    statistics; cuBLAS for the small ones), then one normalise+activate kernel.  Backward: reductions, apply, the input
    gradient (gemm_rows) and the weight gradient (gemm_wgrad).  The bias cancels inside train-mode batch norm: it is
    not added and its gradient is exactly zero."""

    @staticmethod
    def forward(ctx, x, w2d, bias, gamma, beta, running_mean, running_var, decay, eps, act, unbiased_moving_var=True):
        cin, cout = w2d.shape
        x2 = _rows_view(x, cin)
        ws, _ = _zeros((2 * cout,), torch.float64, x2.device)
        # x.W on tensor cores (3xTF32); the GEMM epilogue also accumulates the batch-norm statistics when it can
        y, stats_ready = gemm_rows(x2, w2d, w_kn=True, col_stats=ws)
        rows = y.shape[0]
        out = torch.empty_like(y)
        mean = torch.empty((cout,), device=y.device, dtype=torch.float32)
        var = torch.empty_like(mean)
        _call("pasnl_bn_act_fwd", rows, cout, _ptr(y), _ptr(gamma), _ptr(beta), float(eps), int(act), _ptr(out),
              _ptr(mean), _ptr(var), _ptr(bias), _ptr(running_mean), _ptr(running_var), float(decay),
              1 if unbiased_moving_var else 0, 1 if stats_ready else 0, _ptr(ws), _stream())
        if stats_ready or _bn_single_launch(rows, cout):
            _lib.launch_count -= 2      # no memset / statistics nodes (fused in the GEMM, or one cooperative launch)
        ctx.save_for_backward(x2, w2d, y, gamma, beta, mean, var)
        ctx.meta = (float(eps), int(act), x.shape)
        return out.reshape(x.shape[:-1] + (cout,))

    @staticmethod
    def backward(ctx, dout):
        x2, w2d, y, gamma, beta, mean, var = ctx.saved_tensors
        eps, act, xshape = ctx.meta
        rows, cout = y.shape
        dout = dout.reshape(rows, cout)
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(beta)
        ws = _zero_arena.zeros((2 * cout,), torch.float64) if (_zero_arena is not None and _zero_arena.device == y.device) else None
        ws_zeroed = ws is not None
        if ws is None:
            ws = torch.empty((2 * cout,), device=y.device, dtype=torch.float64)
        _call("pasnl_bn_act_bwd", rows, cout, _ptr(y), _ptr(dout), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(var), eps,
              act, _ptr(dy), _ptr(dgamma), _ptr(dbeta), 1 if ws_zeroed else 0, _ptr(ws), _stream())
        if _bn_single_launch(rows, cout):
            _lib.launch_count -= 2      # clear + reduce + apply ran as one cooperative launch
        elif ws_zeroed:
            _lib.launch_count -= 1      # no memset node was enqueued
        dx = gemm_rows(dy, w2d, w_kn=False).reshape(xshape) if ctx.needs_input_grad[0] else None   # dy.W^T: [n][k] = W as stored
        dw = gemm_wgrad(x2, dy) if ctx.needs_input_grad[1] else None
        dbias = _zeros(gamma.shape, torch.float32, gamma.device)[0] if ctx.needs_input_grad[2] else None
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None, None


def _bn_fwd(y, gamma, beta, eps, act, bias, rm, rv, decay, unbiased, stats_ready, ws):
    rows, cout = y.shape
    out = torch.empty_like(y)
    mean = torch.empty((cout,), device=y.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    _call("pasnl_bn_act_fwd", rows, cout, _ptr(y), _ptr(gamma), _ptr(beta), float(eps), int(act), _ptr(out), _ptr(mean),
          _ptr(var), _ptr(bias), _ptr(rm), _ptr(rv), float(decay), 1 if unbiased else 0, 1 if stats_ready else 0, _ptr(ws),
          _stream())
    if stats_ready or _bn_single_launch(rows, cout):
        _lib.launch_count -= 2
    return out, mean, var


def _bn_bwd(y, dout, gamma, beta, mean, var, eps, act):
    rows, cout = y.shape
    dy = torch.empty_like(y)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(beta)
    ws = _zero_arena.zeros((2 * cout,), torch.float64) if (_zero_arena is not None and _zero_arena.device == y.device) else None
    ws_zeroed = ws is not None
    if ws is None:
        ws = torch.empty((2 * cout,), device=y.device, dtype=torch.float64)
    _call("pasnl_bn_act_bwd", rows, cout, _ptr(y), _ptr(dout), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(var), eps, act,
          _ptr(dy), _ptr(dgamma), _ptr(dbeta), 1 if ws_zeroed else 0, _ptr(ws), _stream())
    if _bn_single_launch(rows, cout):
        _lib.launch_count -= 2
    elif ws_zeroed:
        _lib.launch_count -= 1
    return dy, dgamma, dbeta


def pair_plan(x2, c1, c2):
    """Can conv(cin->c1)+BN+act -> conv(c1->c2)+BN+act on the rows of x2 run with the FIRST activation folded into the
    second layer's kernels?  Needs layer 1's statistics out of its GEMM epilogue, and layer 2's shapes on the
    weights-stationary TMA GEMM (pasnl_gemm_rows_ws) and the MN-major TMA weight gradient (pasnl_gemm_wgrad_tma).
    (Fresh torch allocations are >= 256-byte aligned: the intermediate tensors are represented by that alignment.)"""
    rows, cin = x2.shape
    if _os.environ.get("PASNL_BN_FOLD", "1")[:1] == "0" or rows <= 64:
        return False
    lib = _lib.load()
    ready1 = bool(lib.pasnl_gemm_rows_ws_eligible(rows, cin, c1, x2.data_ptr(), int(x2.stride(0)))) or cin <= GEMM_K_SLAB
    return bool(ready1 and lib.pasnl_gemm_rows_ws_eligible(rows, c1, c2, 256, c1) and c2 % 4 == 0
                and lib.pasnl_gemm_wgrad_tma_eligible(rows, c1, c2, 256, c1, 256, c2))


class DenseBNAct2(torch.autograd.Function):
    """Two stacked 1x1-conv + batch-norm + activation layers (the conv0 -> conv1 pair of a local cell, reference
    pointasnl_util.py:264-270) whose intermediate activation h1 = act(bn(y1)) is never written: layer 1's statistics come
    out of its GEMM epilogue, pasnl_bn_finalize turns them into a per-channel scale / shift, and layer 2's GEMM
    (pasnl_gemm_rows_ws) and weight gradient (pasnl_gemm_wgrad_tma) apply act(y1 * scale + shift) to their operand on
    the way into the tensor core.  Saves one read + one write of the (rows, c1) tensor per step.  When the shapes do not
    qualify (pair_plan) the function runs the two layers exactly as two DenseBNAct calls would.

    raw_out=True (requires pair_plan): layer 2's normalise + activate pass is left to the consumer as well -- returns
    (y2, mean2, var2, scale2, shift2) with y2 the PRE-batch-norm output; the consumer (BNActPointConv) applies
    act(y2 * scale2 + shift2) on chip and returns the gradient with respect to y2, batch-norm terms included."""

    @staticmethod
    def forward(ctx, x, w1, b1, g1, be1, rm1, rv1, w2, b2, g2, be2, rm2, rv2, decay, eps, act1, act2, unbiased,
                raw_out=False):
        cin, c1 = w1.shape
        c2 = w2.shape[1]
        x2 = _rows_view(x, cin)
        fold = pair_plan(x2, c1, c2)
        _require(fold or not raw_out, "DenseBNAct2(raw_out=True) needs shapes that pair_plan() accepts")
        ws1, _ = _zeros((2 * c1,), torch.float64, x2.device)
        y1, ready1 = gemm_rows(x2, w1, w_kn=True, col_stats=ws1)
        rows = y1.shape[0]
        _require(ready1 or not fold, "internal: pair_plan() and gemm_rows() disagree about fused statistics")
        ws2, _ = _zeros((2 * c2,), torch.float64, x2.device)
        dev = y1.device

        def finalize(c, ws, g, be, b, rm, rv):
            mean = torch.empty((c,), device=dev, dtype=torch.float32)
            var, scale, shift = torch.empty_like(mean), torch.empty_like(mean), torch.empty_like(mean)
            _call("pasnl_bn_finalize", rows, c, _ptr(ws), _ptr(g), _ptr(be), float(eps), _ptr(b), _ptr(rm), _ptr(rv),
                  float(decay), 1 if unbiased else 0, _ptr(mean), _ptr(var), _ptr(scale), _ptr(shift), _stream())
            return mean, var, scale, shift
        if fold:
            mean1, var1, scale1, shift1 = finalize(c1, ws1, g1, be1, b1, rm1, rv1)
            y2 = torch.empty((rows, c2), device=dev, dtype=torch.float32)
            _call("pasnl_gemm_rows_ws", rows, c1, c2, _ptr(y1), c1, _ptr(w2), int(w2.stride(0)), 1, _ptr(y2), c2, _ptr(ws2),
                  _ptr(scale1), _ptr(shift1), int(act1), _stream())
            ready2, h1 = True, None
        else:
            h1, mean1, var1 = _bn_fwd(y1, g1, be1, eps, act1, b1, rm1, rv1, decay, unbiased, ready1, ws1)
            scale1 = shift1 = None
            y2, ready2 = gemm_rows(h1, w2, w_kn=True, col_stats=ws2)
        ctx.meta = (float(eps), int(act1), int(act2), x.shape, fold, bool(raw_out))
        if raw_out:
            mean2, var2, scale2, shift2 = finalize(c2, ws2, g2, be2, b2, rm2, rv2)
            ctx.save_for_backward(x2, w1, y1, g1, be1, mean1, var1, scale1, shift1, h1, w2, None, None, None, None, None)
            ctx.mark_non_differentiable(mean2, var2, scale2, shift2)
            return y2.reshape(x.shape[:-1] + (c2,)), mean2, var2, scale2, shift2
        out, mean2, var2 = _bn_fwd(y2, g2, be2, eps, act2, b2, rm2, rv2, decay, unbiased, ready2, ws2)
        ctx.save_for_backward(x2, w1, y1, g1, be1, mean1, var1, scale1, shift1, h1, w2, y2, g2, be2, mean2, var2)
        return out.reshape(x.shape[:-1] + (c2,))

    @staticmethod
    def backward(ctx, dout, *_unused):
        x2, w1, y1, g1, be1, mean1, var1, scale1, shift1, h1, w2, y2, g2, be2, mean2, var2 = ctx.saved_tensors
        eps, act1, act2, xshape, fold, raw_out = ctx.meta
        rows, c1 = y1.shape
        c2 = w2.shape[1]
        dout = dout.reshape(rows, c2)
        if not dout.is_contiguous():
            dout = dout.contiguous()
        if raw_out:
            dy2, dg2, dbe2 = dout, None, None          # the consumer already went back through layer 2's batch norm
        else:
            dy2, dg2, dbe2 = _bn_bwd(y2, dout, g2, be2, mean2, var2, eps, act2)
        dh1 = gemm_rows(dy2, w2, w_kn=False)
        if fold:
            dw2, _zeroed = _zeros((c1, c2), torch.float32, y1.device)
            _call("pasnl_gemm_wgrad_tma", rows, c1, c2, _ptr(y1), c1, _ptr(dy2), c2, _ptr(dw2), c2, 1, _ptr(scale1),
                  _ptr(shift1), act1, _stream())
        else:
            dw2 = gemm_wgrad(h1, dy2)
        dy1, dg1, dbe1 = _bn_bwd(y1, dh1, g1, be1, mean1, var1, eps, act1)
        dx = gemm_rows(dy1, w1, w_kn=False).reshape(xshape) if ctx.needs_input_grad[0] else None
        dw1 = gemm_wgrad(x2, dy1)
        db1 = _zeros(g1.shape, torch.float32, g1.device)[0] if ctx.needs_input_grad[2] else None
        db2 = _zeros((c2,), torch.float32, g1.device)[0] if ctx.needs_input_grad[8] else None
        return (dx, dw1, db1, dg1, dbe1, None, None, dw2, db2, dg2, dbe2, None, None, None, None, None, None, None, None)


class BNActPointConv(torch.autograd.Function):
    """PointConvContract on x = act(batch_norm(y)) with y the PRE-batch-norm output of the producing layer
    (DenseBNAct2(raw_out=True)): the normalise + activate runs on the tile in shared memory (pasnl_pointconv_*_bn), the
    backward continues through the batch norm (pasnl_bn_act_bwd) and returns the gradient with respect to y, gamma, beta."""

    @staticmethod
    def forward(ctx, y, gamma, beta, mean, var, scale, shift, w, eps, act):
        y, w = _f32(y, "y"), _f32(w, "w")
        k, c = y.shape[-2], y.shape[-1]
        _require(w.shape[-1] == 32 and w.shape[:-1] == y.shape[:-1], "weight must be (..., k, 32) matching y (..., k, c)")
        groups = y.numel() // (k * c)
        out = torch.empty(y.shape[:-2] + (c, 32), device=y.device, dtype=torch.float32)
        _call("pasnl_pointconv_fwd_bn", groups, k, c, _ptr(y), _ptr(w), _ptr(out), _ptr(scale), _ptr(shift), int(act),
              _stream())
        ctx.save_for_backward(y, w, gamma, beta, mean, var, scale, shift)
        ctx.meta = (float(eps), int(act))
        return out

    @staticmethod
    def backward(ctx, g):
        y, w, gamma, beta, mean, var, scale, shift = ctx.saved_tensors
        eps, act = ctx.meta
        g = _f32(g, "grad_out")
        k, c = y.shape[-2], y.shape[-1]
        gh, gw = torch.empty_like(y), torch.empty_like(w)
        _call("pasnl_pointconv_bwd_bn", y.numel() // (k * c), k, c, _ptr(y), _ptr(w), _ptr(g), _ptr(gh), _ptr(gw),
              _ptr(scale), _ptr(shift), act, _stream())
        dy, dgamma, dbeta = _bn_bwd(y.reshape(-1, c), gh.reshape(-1, c), gamma, beta, mean, var, eps, act)
        return dy.reshape(y.shape), dgamma, dbeta, None, None, None, None, gw, None, None


def pointconv_eligible(k, c):
    return k % 4 == 0 and c % 4 == 0 and k <= 256 and max(2 * (k * c + k * 32), k * c + k * 32 + 2 * c * 32) * 4 <= 200 * 1024


# ------------------------------------------------------------------ launch-bound glue (csrc/small_ops.cu)
class AddN(torch.autograd.Function):
    """a + b (+ c): the tf.add chains of the layer functions (pointasnl_util.py:283-285) as one kernel."""

    @staticmethod
    def forward(ctx, a, b, c=None):
        a, b = _f32(a, "a"), _f32(b, "b")
        _require(a.shape == b.shape and (c is None or c.shape == a.shape), "add_n expects equal shapes")
        c = _f32(c, "c") if c is not None else None
        out = torch.empty_like(a)
        _call("pasnl_add_n", a.numel(), _ptr(a), _ptr(b), _ptr(c), _ptr(out), _stream())
        ctx.has_c = c is not None
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g, (g if ctx.has_c else None)


def add_n(a, b, c=None):
    if a.is_cuda:
        return AddN.apply(a, b, c)
    return a + b if c is None else a + b + c


class ConcatCols(torch.autograd.Function):
    """torch.cat([a, b], dim=-1) for tensors with equal leading shapes (tf.concat(axis=-1))."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32(a, "a"), _f32(b, "b")
        _require(a.shape[:-1] == b.shape[:-1], "concat_cols expects equal leading shapes")
        ca, cb = a.shape[-1], b.shape[-1]
        rows = a.numel() // ca
        out = torch.empty(a.shape[:-1] + (ca + cb,), device=a.device, dtype=torch.float32)
        _call("pasnl_concat_cols", rows, ca, cb, _ptr(a), ca, _ptr(b), cb, _ptr(out), _stream())
        ctx.dims = (rows, ca, cb, a.shape, b.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        rows, ca, cb, sa, sb = ctx.dims
        g = _f32(g, "grad")
        ga = torch.empty(sa, device=g.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        gb = torch.empty(sb, device=g.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        _call("pasnl_split_cols", rows, ca, cb, _ptr(g), _ptr(ga), _ptr(gb), _stream())
        return ga, gb


def concat_cols(a, b):
    if a.is_cuda:
        return ConcatCols.apply(a, b)
    return torch.cat([a, b], dim=-1)


def split_cols(x, ca):
    """x (..., ca+cb) -> contiguous (..., ca), (..., cb) copies (tf.slice of the input placeholder); no gradient."""
    x = _f32(x.detach(), "x")
    w = x.shape[-1]
    cb = w - ca
    rows = x.numel() // w
    a = torch.empty(x.shape[:-1] + (ca,), device=x.device, dtype=torch.float32)
    b = torch.empty(x.shape[:-1] + (cb,), device=x.device, dtype=torch.float32)
    _call("pasnl_split_cols", rows, ca, cb, _ptr(x), _ptr(a), _ptr(b), _stream())
    return a, b


_dropout_step = None     # device int64 scalar: the optimizer's step counter (set by TrainStep) -> fresh masks per replay
_dropout_calls = 0


def set_dropout_step(t):
    global _dropout_step
    _dropout_step = t


class Dropout(torch.autograd.Function):
    """tf.nn.dropout(x, keep_prob): the mask is a counter-based function of (seed, step, index), regenerated in the
    backward pass instead of being stored."""

    @staticmethod
    def forward(ctx, x, keep_prob, seed):
        x = _f32(x, "x")
        out = torch.empty_like(x)
        _call("pasnl_dropout", x.numel(), _ptr(x), float(keep_prob), int(seed), _ptr(_dropout_step), _ptr(out), _stream())
        ctx.args = (float(keep_prob), int(seed), _dropout_step)
        return out

    @staticmethod
    def backward(ctx, g):
        keep, seed, step = ctx.args
        g = _f32(g, "grad")
        out = torch.empty_like(g)
        _call("pasnl_dropout", g.numel(), _ptr(g), keep, seed, _ptr(step), _ptr(out), _stream())
        return out, None, None


def dropout(x, keep_prob, seed):
    return Dropout.apply(x, keep_prob, seed)


class SoftmaxCE(torch.autograd.Function):
    """Sparse soft-max cross entropy -> scalar.  weights None: out_scale * mean (pointasnl_cls.py:60-61); else
    out_scale * sum(w * ce) / max(count_nonzero(w), 1) (tf.losses.sparse_softmax_cross_entropy, sem_seg_res.py:81)."""

    @staticmethod
    def forward(ctx, logits, labels, weights, out_scale):
        logits = _f32(logits, "logits")
        c = logits.shape[-1]
        rows = logits.numel() // c
        _require(labels.is_cuda and labels.dtype == torch.int64 and labels.numel() == rows, "labels must be CUDA int64 (rows,)")
        labels = labels.contiguous()
        weights = _f32(weights, "weights") if weights is not None else None
        lse = torch.empty((rows,), device=logits.device, dtype=torch.float32)
        ws = torch.empty((4,), device=logits.device, dtype=torch.float64)
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        _call("pasnl_softmax_ce_fwd", rows, c, _ptr(logits), _ptr(labels), _ptr(weights), float(out_scale), _ptr(lse),
              _ptr(ws), _ptr(loss), _stream())
        ctx.save_for_backward(logits, labels, weights, lse, ws)
        ctx.out_scale = float(out_scale)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, labels, weights, lse, ws = ctx.saved_tensors
        c = logits.shape[-1]
        rows = logits.numel() // c
        g = _f32(g, "grad").reshape(())
        d = torch.empty_like(logits)
        _call("pasnl_softmax_ce_bwd", rows, c, _ptr(logits), _ptr(labels), _ptr(weights), ctx.out_scale, _ptr(lse), _ptr(ws),
              _ptr(g), _ptr(d), _stream())
        return d, None, None, None


def softmax_ce(logits, labels, weights=None, out_scale=1.0):
    return SoftmaxCE.apply(logits, labels, weights, out_scale)


class DenseBias(torch.autograd.Function):
    """x @ w + bias for a layer without batch norm (tf_util.fully_connected / conv1d with bn=False: the logits layers)."""

    @staticmethod
    def forward(ctx, x, w2d, bias):
        cin, cout = w2d.shape
        x2 = _rows_view(_f32(x, "x"), cin)
        rows = x2.shape[0]
        if rows <= 64:          # bias-initialised accumulator, then the split-k skinny GEMM adds into it
            out = torch.empty((rows, cout), device=x.device, dtype=torch.float32)
            _call("pasnl_bias_add", rows, cout, None, _ptr(bias), _ptr(out), _stream())
            w = w2d if w2d.stride(1) == 1 else w2d.contiguous()
            _call("pasnl_gemm_skinny", rows, cin, cout, _ptr(x2), int(x2.stride(0)), _ptr(w), int(w.stride(0)), 0,
                  _ptr(out), cout, _stream())
        else:
            y = gemm_rows(x2, w2d, w_kn=True)
            out = torch.empty_like(y)
            _call("pasnl_bias_add", rows, cout, _ptr(y), _ptr(bias), _ptr(out), _stream())
        ctx.save_for_backward(x2, w2d)
        ctx.xshape = x.shape
        return out.reshape(x.shape[:-1] + (cout,))

    @staticmethod
    def backward(ctx, g):
        x2, w2d = ctx.saved_tensors
        cin, cout = w2d.shape
        g2 = _f32(g, "grad").reshape(-1, cout)
        dx = gemm_rows(g2, w2d, w_kn=False).reshape(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = gemm_wgrad(x2, g2) if ctx.needs_input_grad[1] else None
        db = None
        if ctx.needs_input_grad[2]:
            db = torch.empty((cout,), device=g2.device, dtype=torch.float32)
            _call("pasnl_col_sum", g2.shape[0], cout, _ptr(g2), _ptr(db), _stream())
        return dx, dw, db


class SAFront(torch.autograd.Function):
    """The three readers of the grouped tensor new_point (B, S, K, W) = [xyz_j - centre | xyz_j | feature_j] in one
    autograd node (utils/pointasnl_util.py:258,264-268): the skip connection's max over the K neighbours, the identity
    that feeds conv0, and the grouped_xyz = new_point[..., :3] slice that feeds the weight net.
    Backward: the dense gradient arriving through the identity output (conv0's input gradient) is taken as the result
    buffer; the max-pool gradient is SCATTERED into it through the tie mask saved by the forward pass and the 3-column
    slice gradient is added in place -- instead of a dense (B,S,K,W) max-pool gradient, a zero-filled slice gradient
    and two full-tensor adds (cls layer 2: five extra passes over a 140 MB tensor)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "new_point")
        b, s, k, w = x.shape
        _require(k <= 64, "SAFront supports up to 64 neighbours")
        pooled = torch.empty((b, s, w), device=x.device, dtype=torch.float32)
        mask = torch.empty((b, s, w), device=x.device, dtype=torch.int64)
        _call("pasnl_maxpool_k_mask", b * s, k, w, _ptr(x), _ptr(pooled), _ptr(mask), _stream())
        xyz = torch.empty((b, s, k, 3), device=x.device, dtype=torch.float32)
        _call("pasnl_take_cols", b * s * k, 3, _ptr(x), w, _ptr(xyz), _stream())
        ctx.save_for_backward(mask)
        ctx.dims = (b, s, k, w)
        ident = x.view_as(x)
        return pooled, ident, xyz

    @staticmethod
    def backward(ctx, g_pooled, g_ident, g_xyz):
        (mask,) = ctx.saved_tensors
        b, s, k, w = ctx.dims
        if g_ident is None:
            g = torch.zeros((b, s, k, w), device=mask.device, dtype=torch.float32)
        else:
            g = _f32(g_ident, "grad")          # owned by this node: the only consumer of `ident` is conv0
        if g_pooled is not None:
            _call("pasnl_maxpool_k_scatter", b * s, k, w, _ptr(mask), _ptr(_f32(g_pooled, "grad")), _ptr(g), _stream())
        if g_xyz is not None:
            _call("pasnl_add_cols", b * s * k, 3, _ptr(_f32(g_xyz, "grad")), _ptr(g), w, _stream())
        return g


def adam_l2(flat_p, flat_g, m, v, l2_mask, l2_coeff, grad_scale, lr, step, reg_loss, loss_in=None, loss_out=None,
            beta1=0.9, beta2=0.999, eps=1e-8):
    """One TF-style Adam step with the L2 regulariser folded in (pasnl_adam_l2); publishes loss_out = loss_in + reg."""
    _call("pasnl_adam_l2", flat_p.numel(), _ptr(flat_p), _ptr(flat_g), _ptr(m), _ptr(v), _ptr(l2_mask), float(l2_coeff),
          float(grad_scale), float(lr), float(beta1), float(beta2), float(eps), _ptr(step), _ptr(reg_loss), _ptr(loss_in),
          _ptr(loss_out), _stream())
