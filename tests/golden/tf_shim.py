"""A stand-in `tensorflow` module that lets the REFERENCE's own Python files run here, unmodified.

TEST INFRASTRUCTURE (like everything under oracle/).  TensorFlow 1.13 cannot be installed in this image (no network,
no cp312 wheel), but the reference's layer code -- utils/pointasnl_util.py, utils/pointnet_util.py, utils/tf_util.py,
models/pointasnl_cls.py, models/pointasnl_sem_seg_res.py and the three tf_ops/*/tf_*.py wrappers -- is plain Python
over a few dozen `tf.*` calls.  This file implements exactly those calls eagerly on torch-CPU tensors (fp64 by
default), with TF 1.13's documented semantics, so that `import pointasnl_util` imports the file that lies under
/root/reference and `PointASNLSetAbstraction(...)` executes the reference's own statements.  torch autograd supplies
`tf.gradients`.  tests/golden/make_golden_tf.py uses it to emit tests/golden/tf_graph_*.npz; tests/test_oracle.py
pins oracle/model_ref.py (and the GPU tests pin the CUDA path) to those vectors.

What is NOT the reference's code here, and where it comes from instead:
  * `tf.load_op_library(...)` returns op modules whose kernels are the reference's own C++ where it has a CPU
    implementation (three_nn / three_interpolate(+grad): oracle/_ref/ref_interp.so, compiled untouched from
    tf_interpolate.cpp:60-153) and oracle/cpu_ref.c elsewhere (FPS, gather_point, query_ball_point, group_point:
    GPU-only in the reference; cpu_ref.c is pinned bit-exact against the compiled reference kernels on the B200,
    tests/golden/gpu_ref_*.npz).  Gradients go through the functions the reference registers with
    `@tf.RegisterGradient` (tf_sampling.py:43-47, tf_grouping.py:42-46, tf_interpolate.py:29-34).
  * `nearest_neighbors.lib.python.nearest_neighbors` (a Cython shim, knn.pyx:71-109) is replaced by a module whose
    `knn_batch` calls the reference's own `cpp_knn_batch_omp` (knn_.cxx:104-135 + nanoflann, oracle/_ref/ref_knn.so).
  * TF built-ins follow TF 1.13: conv = cross-correlation NHWC/HWIO; contrib batch_norm eps 1e-3, biased batch
    variance for normalisation, moving_variance fed the Bessel-corrected variance on the fused (rank 2/4) path;
    `tf.losses.sparse_softmax_cross_entropy` = weighted sum / number of non-zero weights AND registers its result
    in the 'losses' collection (GraphKeys.LOSSES) -- which is how models/pointasnl_sem_seg_res.py:83 ends up
    counting the classification loss twice; dropout scales by 1/keep_prob (mask from a seeded torch generator, or
    the identity when the shim is installed with dropout=False so that other implementations can be compared).
"""
import contextlib
import os
import sys
import types
import zlib

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE = os.environ.get("PASNL_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


# ------------------------------------------------------------------------------------------------ tensors
class Dim(int):
    """tf.Dimension: an int with `.value`."""

    @property
    def value(self):
        return int(self)


class TensorShape(list):
    def as_list(self):
        return [int(d) for d in self]

    def with_rank(self, r):
        assert len(self) == r
        return self

    @property
    def dims(self):
        return list(self)


class T(torch.Tensor):
    """torch tensor with the handful of tf.Tensor methods the reference calls.  `-=` / `+=` are out-of-place, as in a
    TF graph (pointasnl_util.py:248 rebinds `grouped_xyz`; the tensor concatenated at :74 keeps absolute xyz)."""

    def get_shape(self):
        return TensorShape(Dim(s) for s in self.shape)

    def set_shape(self, shape):
        assert [int(s) for s in shape] == list(self.shape), (shape, self.shape)

    def __isub__(self, other):
        return self - other

    def __iadd__(self, other):
        return self + other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def eval(self, **kw):
        return self.detach().numpy()

    @property
    def name(self):                      # tf.Variable.name ("scope/weights:0"); torch's own `.name` is read-only
        return getattr(self, "_tf_name", None)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if dtype is None:
            dtype = STATE.float_dtype if a.dtype.kind == "f" else (torch.bool if a.dtype.kind == "b" else torch.int64)
        t = torch.as_tensor(a).to(dtype)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.as_subclass(T)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


class _DType:
    def __init__(self, name):
        self.name = name

    def torch(self):
        return {"float32": STATE.float_dtype, "float16": STATE.float_dtype, "float64": torch.float64,
                "int32": torch.int64, "int64": torch.int64, "bool": torch.bool}[self.name]


# ------------------------------------------------------------------------------------------------ graph state
class _State:
    def __init__(self):
        self.reset()

    def reset(self, float_dtype=torch.float64, dropout=False, seed=0, init="golden"):
        self.float_dtype = float_dtype
        self.dropout = dropout
        self.gen = torch.Generator().manual_seed(seed)
        self.scope = []
        self.variables = {}          # TF name (without ':0') -> T
        self.trainable = {}          # name -> bool
        self.collections = {}
        self.gradient_fns = {}       # op type -> python gradient function registered by the reference
        self.init = init             # "golden": golden_init(name, shape, kind);  "tf": xavier / constants


STATE = _State()


from oracle.golden_init import golden_init  # noqa: E402  (shared, name-seeded variable values)


class _Initializer:
    def __init__(self, kind, value=0.0, stddev=1.0):
        self.kind, self.value, self.stddev = kind, value, stddev

    def __call__(self, shape):
        shape = tuple(int(s) for s in shape)
        if self.kind == "constant":
            return torch.full(shape, float(self.value), dtype=torch.float64)
        if self.kind == "xavier":
            fan_in = int(np.prod(shape[:-1]))
            fan_out = int(np.prod(shape[:-2])) * shape[-1] if len(shape) > 2 else shape[-1]
            limit = float(np.sqrt(6.0 / (fan_in + fan_out)))
            return (torch.rand(shape, generator=STATE.gen, dtype=torch.float64) * 2 - 1) * limit
        if self.kind == "truncated_normal":
            t = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(t, 0.0, self.stddev, -2 * self.stddev, 2 * self.stddev, generator=STATE.gen)
            return t
        raise KeyError(self.kind)


def _full_name(name):
    return "/".join(STATE.scope + [name])


def _get_variable(name, shape=None, initializer=None, dtype=None, trainable=True, **kw):
    full = _full_name(name)
    if full in STATE.variables:                      # reuse (AUTO_REUSE), e.g. a second tower / an evaluation pass
        return STATE.variables[full]
    leaf = name.split("/")[-1]
    if STATE.init == "golden" and leaf in ("weights", "biases", "gamma", "beta", "moving_mean", "moving_variance"):
        val = torch.as_tensor(golden_init(full, shape, leaf))
    else:
        val = initializer(shape)
    v = val.to(STATE.float_dtype).as_subclass(T)
    v.requires_grad_(bool(trainable))
    v._tf_name = full + ":0"
    STATE.variables[full] = v
    STATE.trainable[full] = bool(trainable)
    return v


@contextlib.contextmanager
def _variable_scope(name_or_scope, reuse=None, **kw):
    if isinstance(name_or_scope, _ScopeHandle):       # tf.variable_scope(tf.get_variable_scope(), reuse=...)
        yield name_or_scope
        return
    STATE.scope.append(str(name_or_scope))
    try:
        yield _ScopeHandle("/".join(STATE.scope))
    finally:
        STATE.scope.pop()


class _ScopeHandle:
    def __init__(self, name):
        self.name = name


# ------------------------------------------------------------------------------------------------ custom ops
class _Op:
    def __init__(self, type_, inputs, outputs):
        self.type, self.inputs, self.outputs = type_, inputs, outputs


def _np32(t):
    return np.ascontiguousarray(t.detach().numpy(), dtype=np.float32)


def _npi(t):
    return np.ascontiguousarray(t.detach().numpy(), dtype=np.int32)


class _RegisteredGradOp(torch.autograd.Function):
    """Forward = the native kernel; backward = the python function the reference registered with
    @tf.RegisterGradient for this op type, called with a TF-style `op` (inputs / outputs) and the output gradient."""

    @staticmethod
    def forward(ctx, op_type, kernel, *inputs):
        out = kernel(*inputs)
        ctx.op_type = op_type
        ctx.inputs = inputs
        ctx.out = out
        return out

    @staticmethod
    def backward(ctx, grad_out):
        fn = STATE.gradient_fns[ctx.op_type]
        op = _Op(ctx.op_type, [i.as_subclass(T) if isinstance(i, torch.Tensor) else i for i in ctx.inputs],
                 [ctx.out.as_subclass(T)])
        grads = fn(op, grad_out.as_subclass(T))
        return (None, None) + tuple(None if g is None else g.to(grad_out.dtype) for g in grads)


def _float_out(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(STATE.float_dtype).as_subclass(T)


def _int_out(a):
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int64)).as_subclass(T)


def _sampling_module():
    from oracle import cref
    m = types.SimpleNamespace()

    def farthest_point_sample(inp, npoint):                     # tf_sampling.cpp:95-123 (GPU only) -> cpu_ref.c
        return _int_out(cref.fps(_np32(inp), int(npoint)))

    def gather_point(inp, idx):                                 # tf_sampling.cpp:126-148
        k = lambda a, i: _float_out(cref.gather_point(_np32(a), _npi(i)))
        return _RegisteredGradOp.apply("GatherPoint", k, inp, idx)

    def gather_point_grad(inp, idx, out_g):                     # tf_sampling.cpp:151-178
        return _float_out(cref.gather_point_grad(inp.shape[1], _npi(idx), _np32(out_g)))

    def prob_sample(inp, inpr):
        raise NotImplementedError("prob_sample: only the reference file's own demo calls it")

    m.farthest_point_sample, m.gather_point, m.gather_point_grad, m.prob_sample = (
        farthest_point_sample, gather_point, gather_point_grad, prob_sample)
    return m


def _grouping_module():
    from oracle import cref
    m = types.SimpleNamespace()

    def query_ball_point(xyz1, xyz2, radius, nsample):          # tf_grouping.cpp:67-106 (GPU only) -> cpu_ref.c
        idx, cnt = cref.ball_query(float(radius), int(nsample), _np32(xyz1), _np32(xyz2))
        return _int_out(idx), _int_out(cnt)

    def group_point(points, idx):                               # tf_grouping.cpp:142-171
        k = lambda p, i: _float_out(cref.group_point(_np32(p), _npi(i)))
        return _RegisteredGradOp.apply("GroupPoint", k, points, idx)

    def group_point_grad(points, idx, grad_out):                # tf_grouping.cpp:174-208
        return _float_out(cref.group_point_grad(points.shape[1], _npi(idx), _np32(grad_out)))

    def selection_sort(dist, k):                                # tf_grouping.cpp:109-139
        outi, out = cref.selection_sort(int(k), _np32(dist))
        return _int_out(outi), _float_out(out)

    m.query_ball_point, m.group_point, m.group_point_grad, m.selection_sort = (
        query_ball_point, group_point, group_point_grad, selection_sort)
    return m


def _interpolate_module():
    from oracle import cref, refnative
    impl = refnative if refnative.available("ref_interp.so") else cref   # the reference's own CPU code when built
    m = types.SimpleNamespace()

    def three_nn(xyz1, xyz2):                                   # tf_interpolate.cpp:60-103
        d, i = impl.three_nn(_np32(xyz1), _np32(xyz2))
        return _float_out(d), _int_out(i)

    def three_interpolate(points, idx, weight):                 # tf_interpolate.cpp:107-129
        k = lambda p, i, w: _float_out(impl.three_interpolate(_np32(p), _npi(i), _np32(w)))
        return _RegisteredGradOp.apply("ThreeInterpolate", k, points, idx, weight)

    def three_interpolate_grad(points, idx, weight, grad_out):  # tf_interpolate.cpp:131-153
        return _float_out(impl.three_interpolate_grad(points.shape[1], _npi(idx), _np32(weight), _np32(grad_out)))

    m.three_nn, m.three_interpolate, m.three_interpolate_grad = three_nn, three_interpolate, three_interpolate_grad
    return m


def _load_op_library(path):
    base = os.path.basename(path)
    if "sampling" in base:
        return _sampling_module()
    if "grouping" in base:
        return _grouping_module()
    if "interpolate" in base:
        return _interpolate_module()
    raise FileNotFoundError(path)


def _register_gradient(op_type):
    def deco(fn):
        STATE.gradient_fns[op_type] = fn
        return fn
    return deco


# ------------------------------------------------------------------------------------------------ tf.* functions
def _gather_nd(params, indices):
    nd = indices.shape[-1]
    return params[tuple(indices[..., i] for i in range(nd))]


def _slice(t, begin, size):
    sl = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
    return t[sl]


def _reduce(fn):
    def red(x, axis=None, keepdims=False, keep_dims=None, name=None, reduction_indices=None):
        if keep_dims is not None:
            keepdims = keep_dims
        if axis is None:
            axis = reduction_indices
        x = _t(x)
        ax = _axes(axis)
        if ax is None:
            ax = tuple(range(x.dim()))
        return fn(x, ax, bool(keepdims))
    return red


def _amax(x, ax, keep):
    # tf.reduce_max's gradient is split evenly among tied maxima -- torch.amax does the same
    return torch.amax(x, dim=ax, keepdim=keep)


def _matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def _conv2d(inputs, kernel, strides, padding, data_format="NHWC", **kw):
    """tf.nn.conv2d: NHWC input, HWIO filter, cross-correlation."""
    assert data_format == "NHWC"
    kh, kw_, cin, cout = kernel.shape
    x = inputs.permute(0, 3, 1, 2)
    w = kernel.permute(3, 2, 0, 1)
    if padding == "SAME":
        assert kh % 2 == 1 and kw_ % 2 == 1
        pad = (kh // 2, kw_ // 2)
    else:
        pad = 0
    y = F.conv2d(x, w, None, stride=(int(strides[1]), int(strides[2])), padding=pad)
    return y.permute(0, 2, 3, 1)


def _conv1d(inputs, kernel, stride, padding, data_format="NHWC", **kw):
    assert data_format == "NHWC"
    k, cin, cout = kernel.shape
    pad = k // 2 if padding == "SAME" else 0
    y = F.conv1d(inputs.permute(0, 2, 1), kernel.permute(2, 1, 0), None, stride=int(stride), padding=pad)
    return y.permute(0, 2, 1)


def _as_bool(pred):
    if isinstance(pred, torch.Tensor):
        return bool(pred.item())
    return bool(pred)


def _batch_norm(inputs, decay=0.999, center=False, scale=False, epsilon=0.001, is_training=True,
                updates_collections="update_ops", scope=None, data_format="NHWC", fused=None, **kw):
    """tf.contrib.layers.batch_norm (TF 1.13 contrib/layers/python/layers/layers.py).  Statistics over every axis but
    the last; normalisation with the biased batch variance.  updates_collections=None -> the moving averages are
    updated in place as part of the forward pass: moving = decay*moving + (1-decay)*batch.  Rank-2 / rank-4 inputs
    take the fused kernel, which feeds the moving variance the Bessel-corrected batch variance; other ranks the
    biased one."""
    assert data_format == "NHWC"
    c = inputs.shape[-1]
    with _variable_scope(scope if scope is not None else "BatchNorm"):
        beta = _get_variable("beta", (c,), _Initializer("constant", 0.0)) if center else None
        gamma = _get_variable("gamma", (c,), _Initializer("constant", 1.0)) if scale else None
        mmean = _get_variable("moving_mean", (c,), _Initializer("constant", 0.0), trainable=False)
        mvar = _get_variable("moving_variance", (c,), _Initializer("constant", 1.0), trainable=False)
    dims = tuple(range(inputs.dim() - 1))
    if _as_bool(is_training):
        mean = inputs.mean(dim=dims)
        var = ((inputs - mean) ** 2).mean(dim=dims)
        count = inputs.numel() // c
        use_fused = inputs.dim() in (2, 4) if fused is None else bool(fused)
        upd_var = var * (count / max(count - 1, 1)) if use_fused else var
        d = float(decay)
        with torch.no_grad():
            mmean.mul_(d).add_((1 - d) * mean.detach())
            mvar.mul_(d).add_((1 - d) * upd_var.detach())
    else:
        mean, var = mmean, mvar
    y = (inputs - mean) * torch.rsqrt(var + epsilon)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y


def _dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
    if not STATE.dropout:
        return x
    shape = tuple(x.shape) if noise_shape is None else tuple(int(s) for s in noise_shape)
    mask = (torch.rand(shape, generator=STATE.gen, dtype=torch.float64) < float(keep_prob)).to(x.dtype)
    return x * mask / float(keep_prob)


def _sparse_ce(logits=None, labels=None, name=None, **kw):
    lg = logits.reshape(-1, logits.shape[-1])
    ce = F.cross_entropy(lg, _t(labels).reshape(-1).long(), reduction="none")
    return ce.reshape(logits.shape[:-1])


def _losses_sparse_ce(labels, logits, weights=1.0, scope=None, loss_collection="losses", reduction=None):
    """tf.losses.sparse_softmax_cross_entropy: Reduction.SUM_BY_NONZERO_WEIGHTS, and the result is added to
    `loss_collection` (default GraphKeys.LOSSES == 'losses')."""
    ce = _sparse_ce(logits=logits, labels=labels)
    w = _t(weights, STATE.float_dtype)
    w = torch.broadcast_to(w, ce.shape)
    nz = torch.count_nonzero(w)
    loss = (ce * w).sum() / nz.clamp(min=1).to(ce.dtype)
    if loss_collection:
        STATE.collections.setdefault(loss_collection, []).append(loss)
    return loss


def _top_k(x, k=1, sorted=True, name=None):
    v, i = torch.topk(x, int(k), dim=-1, largest=True, sorted=True)
    return v, i


def _py_func(func, inp, Tout, stateful=True, name=None):
    args = []
    for a in inp:
        if isinstance(a, torch.Tensor):
            a = a.detach().numpy()
            if a.dtype.kind == "f":
                a = a.astype(np.float32)      # the graph tensors are tf.float32
            else:
                a = a.astype(np.int32)
        args.append(a)
    out = func(*args)
    return _t(np.asarray(out), Tout.torch() if isinstance(Tout, _DType) else None)


def _gradients(ys, xs, grad_ys=None, **kw):
    ys = ys if isinstance(ys, (list, tuple)) else [ys]
    y = sum(t.sum() for t in ys) if grad_ys is None else sum((t * g).sum() for t, g in zip(ys, grad_ys))
    return list(torch.autograd.grad(y, xs, allow_unused=True, retain_graph=True))


class _Noop:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def as_default(self):
        return self

    def run(self, fetches, feed_dict=None):
        return fetches


def build_module():
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "1.13.1-shim"
    for n in ("float32", "float16", "float64", "int32", "int64", "bool"):
        setattr(tf, n, _DType(n))
    tf.Tensor = T
    tf.TensorShape = lambda dims: TensorShape(Dim(d) for d in dims)
    tf.Graph = _Noop
    tf.Session = _Noop
    tf.device = lambda *a, **k: _Noop()
    tf.control_dependencies = lambda *a, **k: _Noop()
    tf.no_op = lambda *a, **k: None
    tf.variable_scope = _variable_scope
    tf.get_variable_scope = lambda: _ScopeHandle("/".join(STATE.scope))
    tf.get_variable = _get_variable
    tf.global_variables = lambda: list(STATE.variables.values())
    tf.trainable_variables = lambda: [v for k, v in STATE.variables.items() if STATE.trainable[k]]
    tf.Variable = lambda v, trainable=True, name=None, **k: _t(v)
    tf.constant_initializer = lambda value=0.0, **k: _Initializer("constant", value)
    tf.truncated_normal_initializer = lambda mean=0.0, stddev=1.0, **k: _Initializer("truncated_normal", stddev=stddev)
    tf.add_to_collection = lambda name, value: STATE.collections.setdefault(name, []).append(value)
    tf.get_collection = lambda name, scope=None: list(STATE.collections.get(name, []))
    tf.load_op_library = _load_op_library
    tf.RegisterGradient = _register_gradient
    tf.RegisterShape = lambda *a, **k: (lambda fn: fn)
    tf.py_func = _py_func
    tf.gradients = _gradients
    tf.placeholder = lambda dtype, shape=None, name=None: torch.zeros(tuple(int(s) for s in shape),
                                                                      dtype=dtype.torch()).as_subclass(T)
    tf.constant = lambda v, dtype=None, shape=None, name=None: _t(v, dtype.torch() if dtype is not None else None)
    tf.zeros = lambda shape, dtype=None, name=None: torch.zeros(tuple(int(s) for s in shape),
                                                                dtype=(dtype or tf.float32).torch()).as_subclass(T)
    tf.ones_like = lambda t, **k: torch.ones_like(_t(t))
    tf.random_uniform = lambda shape, minval=0, maxval=1.0, dtype=None, **k: (
        torch.rand(tuple(int(s) for s in shape), generator=STATE.gen, dtype=torch.float64) * (maxval - minval) + minval
    ).to(STATE.float_dtype).as_subclass(T)
    tf.range = lambda *a, **k: torch.arange(*[int(x) for x in a]).as_subclass(T)
    tf.cast = lambda x, dtype, name=None: _t(x, dtype.torch())
    tf.identity = lambda x, name=None: x
    tf.reshape = lambda t, shape, name=None: _t(t).reshape(tuple(int(s) for s in shape))
    tf.expand_dims = lambda t, axis=None, name=None, dim=None: _t(t).unsqueeze(int(axis if axis is not None else dim))
    tf.squeeze = lambda t, axis=None, name=None, squeeze_dims=None: (
        _squeeze(_t(t), axis if axis is not None else squeeze_dims))
    tf.tile = lambda t, multiples, name=None: _t(t).repeat(tuple(int(m) for m in multiples))
    tf.concat = lambda values, axis, name=None: torch.cat([_t(v) for v in values], dim=int(axis))
    tf.transpose = lambda t, perm=None, name=None: _t(t).permute(tuple(int(p) for p in perm))
    tf.slice = lambda t, begin, size, name=None: _slice(_t(t), begin, size)
    tf.gather_nd = lambda params, indices, name=None: _gather_nd(_t(params), _t(indices))
    tf.matmul = _matmul
    tf.multiply = lambda a, b, name=None: _t(a) * _t(b)
    tf.add = lambda a, b, name=None: _t(a) + _t(b)
    tf.add_n = _add_n
    tf.maximum = lambda a, b, name=None: torch.maximum(*_pair(a, b))
    tf.minimum = lambda a, b, name=None: torch.minimum(*_pair(a, b))
    tf.sqrt = lambda x, name=None: torch.sqrt(_t(x))
    tf.exp = lambda x, name=None: torch.exp(_t(x))
    tf.abs = lambda x, name=None: torch.abs(_t(x))
    tf.norm = lambda x, axis=None, keepdims=False, **k: torch.linalg.vector_norm(_t(x), dim=_axes(axis), keepdim=keepdims)
    tf.cross = lambda a, b, name=None: torch.linalg.cross(_t(a), _t(b), dim=-1)
    tf.reduce_sum = _reduce(lambda x, ax, keep: torch.sum(x, dim=ax, keepdim=keep))
    tf.reduce_mean = _reduce(lambda x, ax, keep: torch.mean(x, dim=ax, keepdim=keep))
    tf.reduce_max = _reduce(_amax)
    tf.cond = lambda pred, true_fn=None, false_fn=None, **k: (true_fn() if _as_bool(pred) else false_fn())

    nn = types.ModuleType("tensorflow.nn")
    nn.relu = lambda x, name=None: torch.relu(x)
    nn.leaky_relu = lambda x, alpha=0.2, name=None: F.leaky_relu(x, alpha)
    nn.sigmoid = lambda x, name=None: torch.sigmoid(x)
    nn.softmax = lambda x, axis=-1, name=None, dim=None: torch.softmax(_t(x), dim=int(axis if dim is None else dim))
    nn.conv2d = _conv2d
    nn.conv1d = _conv1d
    nn.bias_add = lambda x, b, data_format=None, name=None: x + b
    nn.l2_loss = lambda t, name=None: (_t(t) ** 2).sum() / 2
    nn.dropout = _dropout
    nn.top_k = _top_k
    nn.moments = lambda x, axes, **k: (x.mean(dim=tuple(axes)), x.var(dim=tuple(axes), unbiased=False))
    nn.batch_normalization = lambda x, mean, var, offset, scale, eps, name=None: (
        (x - mean) * torch.rsqrt(var + eps) * scale + offset)
    nn.sparse_softmax_cross_entropy_with_logits = _sparse_ce
    tf.nn = nn

    losses = types.ModuleType("tensorflow.losses")
    losses.sparse_softmax_cross_entropy = _losses_sparse_ce
    tf.losses = losses

    summary = types.ModuleType("tensorflow.summary")
    summary.scalar = summary.histogram = lambda *a, **k: None
    tf.summary = summary

    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = lambda *a, **k: _Initializer("xavier")
    layers.batch_norm = _batch_norm
    contrib.layers = layers
    tf.contrib = contrib

    train = types.ModuleType("tensorflow.train")
    tf.train = train

    python = types.ModuleType("tensorflow.python")
    framework = types.ModuleType("tensorflow.python.framework")
    ops = types.ModuleType("tensorflow.python.framework.ops")
    ops.NoGradient = ops.NotDifferentiable = lambda op_type: None
    framework.ops = ops
    python.framework = framework
    tf.python = python
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.losses": losses, "tensorflow.summary": summary,
            "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers, "tensorflow.train": train,
            "tensorflow.python": python, "tensorflow.python.framework": framework,
            "tensorflow.python.framework.ops": ops}
    return mods


def _squeeze(t, axis):
    if axis is None:
        return t.squeeze()
    ax = _axes(axis)
    for a in sorted(ax if isinstance(ax, tuple) else (ax,), reverse=True):
        assert t.shape[a] == 1, (t.shape, a)     # TF raises on squeezing a non-unit axis
        t = t.squeeze(a)
    return t


def _pair(a, b):
    ta, tb = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
    ref = a if ta else b
    a = a if ta else torch.as_tensor(a, dtype=ref.dtype)
    b = b if tb else torch.as_tensor(b, dtype=ref.dtype)
    return a, b


def _add_n(inputs, name=None):
    inputs = list(inputs)
    if not inputs:
        raise ValueError("inputs must be a list of at least one Tensor/IndexedSlices with the same dtype and shape")
    out = inputs[0]
    for t in inputs[1:]:
        out = out + t
    return out


def _knn_module():
    from oracle import refnative
    m = types.ModuleType("nearest_neighbors.lib.python.nearest_neighbors")

    def knn_batch(pts, queries, K, omp=False):
        """knn.pyx:71-109 -> cpp_knn_batch(_omp) (knn_.cxx:104-135), the reference's own kd-tree code."""
        return refnative.knn_batch(pts, queries, int(K))

    m.knn_batch = knn_batch
    return m


def install(float_dtype=torch.float64, dropout=False, seed=0, init="golden"):
    """Put the stand-in modules into sys.modules and the reference's directories on sys.path, reset the graph state
    (variables, collections, scope) and return the fake `tf` module."""
    if not os.path.isdir(REFERENCE):
        raise FileNotFoundError(f"{REFERENCE} is not mounted: the shim executes the reference's own files")
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    fns = STATE.gradient_fns
    STATE.reset(float_dtype, dropout, seed, init)
    STATE.gradient_fns = fns                      # registered at import time of the reference's tf_ops wrappers
    if "tensorflow" not in sys.modules or not getattr(sys.modules["tensorflow"], "__version__", "").endswith("shim"):
        sys.modules.update(build_module())
        knn = _knn_module()
        for name in ("nearest_neighbors", "nearest_neighbors.lib", "nearest_neighbors.lib.python"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["nearest_neighbors.lib.python.nearest_neighbors"] = knn
        sys.modules["nearest_neighbors.lib.python"].nearest_neighbors = knn
        sys.modules["nearest_neighbors.lib"].python = sys.modules["nearest_neighbors.lib.python"]
        sys.modules["nearest_neighbors"].lib = sys.modules["nearest_neighbors.lib"]
        for sub in ("utils", "models", "tf_ops/sampling", "tf_ops/grouping", "tf_ops/3d_interpolation"):
            p = os.path.join(REFERENCE, sub)
            if p not in sys.path:
                sys.path.append(p)
    return sys.modules["tensorflow"]


def variables():
    """TF name -> tensor for every variable created since install()."""
    return dict(STATE.variables)


def tensor(a, dtype=None):
    return _t(a, dtype)
