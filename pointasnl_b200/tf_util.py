"""Host-side mirror of the reference's utils/tf_util.py for the callers of the hot path.

The reference builds its dense glue (1x1 conv + bias + batch norm + activation, fully connected, dropout) as TF1
graph ops that create variables by *scope name* (tf.variable_scope / tf.get_variable).  To keep
``pointasnl_util.PointASNLSetAbstraction(xyz, feature, ..., scope, ...)`` and friends drop-in at the signature
level, this module reproduces exactly that: a ``VariableStore`` (an nn.Module) in which variables are created on
first use under the same nested scope names TF would give them (``layer1/layer1/conv_kv/weights`` ...), and reused
afterwards (AUTO_REUSE).  Call semantics follow the reference: bias -> batch norm -> activation order
(reference tf_util.py:170-185), NHWC layout, xavier init, BN eps 1e-3 with biased batch variance (:512-531).

On CUDA tensors in training mode the sequence matmul -> bias -> batch norm -> activation runs on the library's
tcgen05 3xTF32 GEMMs (csrc/gemm_tc.cu) and fused batch-norm kernels (csrc/dense_glue.cu) through `_ops.DenseBNAct`
(SURVEY.md 8f rank 2).  The op-by-op PyTorch path below remains for CPU tensors (unit tests of this mirror) and
inference mode, and is checked against the fused path on the GPU (tests/test_gpu_glue.py).

Moving averages follow tf.contrib.layers.batch_norm(updates_collections=None): moving = decay*moving +
(1-decay)*batch, with the Bessel-corrected batch variance for conv2d / fully_connected (TF's fused kernel) and the
biased one for conv1d (rank-3 inputs); pinned by tests/golden/tf_graph_*.npz (`moving/...` entries).
"""
import contextlib
import math

import torch
import torch.nn.functional as F

relu = torch.relu


def leaky_relu(x):
    return F.leaky_relu(x, 0.2)  # tf.nn.leaky_relu default alpha


class VariableStore(torch.nn.Module):
    """TF1-style variable store: ``get_variable`` creates a Parameter (or buffer) under the current scope path on
    first use and returns the existing one afterwards."""

    def __init__(self, device="cuda", seed=0, init_fn=None):
        """init_fn(full_name, shape, leaf_name) -> array-like or None: overrides the initialisers (used to load
        name-seeded golden values or an imported checkpoint at creation time)."""
        super().__init__()
        self.init_fn = init_fn
        self.device = torch.device(device)
        self._scope = []
        self.seed = int(seed)
        self._gen = torch.Generator(device="cpu")
        self._gen.manual_seed(seed)
        self.weight_decay = {}   # variable name -> wd  (tf.add_to_collection('losses', wd * l2_loss(var)))
        self._names = {}         # TF name -> attribute key

    # -- scopes
    @contextlib.contextmanager
    def variable_scope(self, name):
        self._scope.append(str(name))
        try:
            yield
        finally:
            self._scope.pop()

    def full_name(self, name):
        return "/".join(self._scope + [name])

    @staticmethod
    def _key(full):
        return full.replace("/", "__").replace(".", "_")

    # -- variables
    def get_variable(self, name, shape, init, trainable=True, wd=None):
        full = self.full_name(name)
        key = self._key(full)
        if full in self._names:
            v = getattr(self, key)
            if tuple(v.shape) != tuple(shape):
                raise ValueError(f"variable {full} exists with shape {tuple(v.shape)}, requested {tuple(shape)}")
            return v
        value = None
        if self.init_fn is not None:
            value = self.init_fn(full, tuple(shape), name)
        if value is None:
            value = init(tuple(shape), self._gen)
        value = torch.as_tensor(value).to(torch.float32).reshape(tuple(shape)).to(self.device)
        if trainable:
            self.register_parameter(key, torch.nn.Parameter(value))
        else:
            self.register_buffer(key, value)
        self._names[full] = key
        if wd is not None:
            self.weight_decay[full] = float(wd)
        return getattr(self, key)

    def variables(self):
        """TF name -> tensor, for every variable created so far."""
        return {full: getattr(self, key) for full, key in self._names.items()}

    def load_variables(self, values):
        with torch.no_grad():
            for full, val in values.items():
                getattr(self, self._names[full]).copy_(torch.as_tensor(val).to(self.device))

    def regularization_loss(self):
        """sum of wd * l2_loss(var) over variables created with weight decay (tf.nn.l2_loss = sum(v^2)/2)."""
        if not self.weight_decay:
            return torch.zeros((), device=self.device)
        names = list(self.weight_decay)
        return l2_loss_sum([getattr(self, self._names[n]) for n in names], [self.weight_decay[n] for n in names])


class _L2LossSum(torch.autograd.Function):
    """sum_i c_i * tf.nn.l2_loss(v_i) over a LIST of variables in a handful of multi-tensor kernels (the op-by-op
    form is pow/sum/mul forward and two muls backward PER VARIABLE: ~150 two-microsecond kernels a step)."""

    @staticmethod
    def forward(ctx, coeffs, *tensors):
        ctx.coeffs = coeffs
        ctx.save_for_backward(*tensors)
        norms = torch.stack(torch._foreach_norm(list(tensors)))
        return 0.5 * torch.dot(norms * norms, _coeff_vector(coeffs, norms.device))

    @staticmethod
    def backward(ctx, g):
        grads = torch._foreach_mul(list(ctx.saved_tensors), list(ctx.coeffs))
        torch._foreach_mul_(grads, g)
        return (None, *grads)


_coeff_cache = {}


def _coeff_vector(coeffs, device):
    """Device copy of a coefficient tuple, made once (a host->device copy is not allowed during graph capture)."""
    key = (coeffs, str(device))
    if key not in _coeff_cache:
        _coeff_cache[key] = torch.tensor(coeffs, dtype=torch.float32, device=device)
    return _coeff_cache[key]


def l2_loss_sum(tensors, coeffs=None):
    """sum_i coeffs[i] * sum(tensors[i]**2) / 2."""
    tensors = list(tensors)
    coeffs = tuple(float(c) for c in (coeffs if coeffs is not None else [1.0] * len(tensors)))
    return _L2LossSum.apply(coeffs, *tensors)


_default_store = None


def get_default_store():
    if _default_store is None:
        raise RuntimeError("no VariableStore is active: wrap the model call in `with store.as_default():`")
    return _default_store


@contextlib.contextmanager
def as_default(store):
    global _default_store
    prev = _default_store
    _default_store = store
    try:
        yield store
    finally:
        _default_store = prev


VariableStore.as_default = lambda self: as_default(self)


def variable_scope(name):
    return get_default_store().variable_scope(name)


# ---------------------------------------------------------------- initialisers
def xavier_init(fan_in, fan_out):
    limit = math.sqrt(6.0 / (fan_in + fan_out))  # tf.contrib.layers.xavier_initializer (uniform)

    def init(shape, gen):
        return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * limit
    return init


def truncated_normal_init(stddev):
    def init(shape, gen):
        t = torch.empty(shape)
        torch.nn.init.trunc_normal_(t, 0.0, stddev, -2 * stddev, 2 * stddev, generator=gen)
        return t
    return init


def constant_init(value):
    def init(shape, gen):
        return torch.full(shape, float(value))
    return init


def _weights(shape, fan_in, fan_out, use_xavier, stddev, wd):
    init = xavier_init(fan_in, fan_out) if use_xavier else truncated_normal_init(stddev)
    return get_default_store().get_variable("weights", shape, init, wd=wd)


# ---------------------------------------------------------------- fused GEMM + batch norm + activation (CUDA, training)
def _act_code(activation_fn):
    if activation_fn is None:
        return 0
    if activation_fn is relu:
        return 1
    if activation_fn is leaky_relu:
        return 2
    return None


def _fused_dense_bn(x, w2d, bias, scope, bn_decay, is_training, activation_fn, fused_kernel=True):
    """conv/fc + bias + batch_norm(train) + activation through the fused pasnl epilogue; None if not applicable
    (CPU tensors, inference mode, exotic activation): the caller then takes the op-by-op path."""
    code = _act_code(activation_fn)
    if not (x.is_cuda and bool(is_training) and code is not None and x.dtype == torch.float32):
        return None
    from . import _ops
    store = get_default_store()
    c = w2d.shape[1]
    decay = 0.9 if bn_decay is None else float(bn_decay)
    with store.variable_scope(scope):
        beta = store.get_variable("beta", (c,), constant_init(0.0))
        gamma = store.get_variable("gamma", (c,), constant_init(1.0))
        mean = store.get_variable("moving_mean", (c,), constant_init(0.0), trainable=False)
        var = store.get_variable("moving_variance", (c,), constant_init(1.0), trainable=False)
    return _ops.DenseBNAct.apply(x, w2d, bias, gamma, beta, mean, var, decay, 1e-3, code, fused_kernel)


# ---------------------------------------------------------------- batch norm
def batch_norm_template(inputs, is_training, scope, moments_dims_unused, bn_decay, data_format="NHWC"):
    """tf.contrib.layers.batch_norm(center=True, scale=True, decay=bn_decay, updates_collections=None), eps 1e-3
    (reference tf_util.py:512-531): statistics over every axis but the channel one."""
    if data_format != "NHWC":
        raise NotImplementedError("only the NHWC layout the hot path uses is mirrored")
    decay = 0.9 if bn_decay is None else float(bn_decay)
    store = get_default_store()
    c = inputs.shape[-1]
    with store.variable_scope(scope):
        beta = store.get_variable("beta", (c,), constant_init(0.0))
        gamma = store.get_variable("gamma", (c,), constant_init(1.0))
        mean = store.get_variable("moving_mean", (c,), constant_init(0.0), trainable=False)
        var = store.get_variable("moving_variance", (c,), constant_init(1.0), trainable=False)
    x2 = inputs.reshape(-1, c)
    if not bool(is_training):
        return F.batch_norm(x2, mean, var, gamma, beta, False, 0.0, 1e-3).reshape(inputs.shape)
    m = x2.mean(dim=0)
    v = x2.var(dim=0, unbiased=False)
    y = (x2 - m) * torch.rsqrt(v + 1e-3) * gamma + beta
    with torch.no_grad():
        # rank-2 / rank-4 inputs take TF's fused kernel, which feeds the moving variance the Bessel-corrected batch
        # variance; rank-3 (conv1d) inputs the biased one (tf.contrib.layers.batch_norm, TF 1.13)
        rows = x2.shape[0]
        upd = v * (rows / max(rows - 1, 1)) if inputs.dim() in (2, 4) else v
        mean.mul_(decay).add_((1.0 - decay) * m)
        var.mul_(decay).add_((1.0 - decay) * upd)
    return y.reshape(inputs.shape)


def batch_norm_for_fc(inputs, is_training, bn_decay, scope):
    return batch_norm_template(inputs, is_training, scope, [0], bn_decay)


def batch_norm_for_conv1d(inputs, is_training, bn_decay, scope, data_format="NHWC"):
    return batch_norm_template(inputs, is_training, scope, [0, 1], bn_decay, data_format)


def batch_norm_for_conv2d(inputs, is_training, bn_decay, scope, data_format="NHWC"):
    return batch_norm_template(inputs, is_training, scope, [0, 1, 2], bn_decay, data_format)


# ---------------------------------------------------------------- layers (signatures as in the reference)
def conv1d(inputs, num_output_channels, kernel_size, scope, stride=1, padding="SAME", data_format="NHWC",
           use_xavier=True, stddev=1e-3, weight_decay=None, activation_fn=relu, bn=False, bn_decay=None,
           is_training=None):
    """inputs (B, L, C).  Only kernel_size 1 / stride 1 occurs on the hot path (reference tf_util.py:60-115)."""
    if kernel_size != 1 or stride != 1 or data_format != "NHWC":
        raise NotImplementedError("conv1d mirror supports kernel_size=1, stride=1, NHWC (all the hot path uses)")
    store = get_default_store()
    cin = inputs.shape[-1]
    with store.variable_scope(scope):
        w = _weights((1, cin, num_output_channels), cin, num_output_channels, use_xavier, stddev, weight_decay)
        b = store.get_variable("biases", (num_output_channels,), constant_init(0.0))
        w2 = w.view(cin, num_output_channels)   # a view, not w[0]: select_backward is a zero-fill + copy per layer
        if bn:
            fused = _fused_dense_bn(inputs, w2, b, "bn", bn_decay, is_training, activation_fn, fused_kernel=False)
            if fused is not None:
                return fused
        if not bn:
            return _dense_bias(inputs, w2, b, activation_fn)
        out = torch.matmul(inputs, w2) + b
        out = batch_norm_for_conv1d(out, is_training, bn_decay, "bn", data_format)
        if activation_fn is not None:
            out = activation_fn(out)
    return out


def conv2d_pair(inputs, channels, scopes, bn_decay=None, is_training=None, weight_decay=None, activation_fn=relu,
                use_xavier=True, stddev=1e-3):
    """Two stacked conv2d([1,1], bn=True) layers -- the same variables, in the same order and scopes, as
    conv2d(conv2d(inputs, channels[0], scope=scopes[0]), channels[1], scope=scopes[1]) -- run as one fused pair on CUDA in
    training mode (_ops.DenseBNAct2: the activation between the two layers is never written to memory); otherwise the
    two conv2d calls themselves."""
    code = _act_code(activation_fn)
    if not (inputs.is_cuda and bool(is_training) and code is not None and inputs.dtype == torch.float32):
        for c, sc in zip(channels, scopes):
            inputs = conv2d(inputs, c, [1, 1], sc, padding="VALID", bn=True, bn_decay=bn_decay, is_training=is_training,
                            weight_decay=weight_decay, activation_fn=activation_fn, use_xavier=use_xavier, stddev=stddev)
        return inputs
    from . import _ops
    store = get_default_store()
    decay = 0.9 if bn_decay is None else float(bn_decay)
    cin = inputs.shape[-1]
    args = []
    for c, sc in zip(channels, scopes):
        with store.variable_scope(sc):
            w = _weights((1, 1, cin, c), cin, c, use_xavier, stddev, weight_decay)
            bias = store.get_variable("biases", (c,), constant_init(0.0))
            with store.variable_scope("bn"):
                beta = store.get_variable("beta", (c,), constant_init(0.0))
                gamma = store.get_variable("gamma", (c,), constant_init(1.0))
                mean = store.get_variable("moving_mean", (c,), constant_init(0.0), trainable=False)
                var = store.get_variable("moving_variance", (c,), constant_init(1.0), trainable=False)
        args += [w.view(cin, c), bias, gamma, beta, mean, var]
        cin = c
    return _ops.DenseBNAct2.apply(inputs, *args, decay, 1e-3, code, code, True)


def conv2d(inputs, num_output_channels, kernel_size, scope, stride=[1, 1], padding="SAME", data_format="NHWC",
           use_xavier=True, stddev=1e-3, weight_decay=None, activation_fn=relu, bn=False, bn_decay=None,
           is_training=None):
    """inputs (B, H, W, C).  Supported kernels: [1,1], and [1,W] with padding='VALID' (the PointConv-style
    `after_conv`, reference pointasnl_util.py:275-278), which is a matmul over the flattened (W, C) axes
    (reference tf_util.py:120-185)."""
    kh, kw = kernel_size
    if kh != 1 or list(stride) != [1, 1] or data_format != "NHWC":
        raise NotImplementedError("conv2d mirror supports kernel height 1, stride 1, NHWC")
    store = get_default_store()
    b_, h, wdt, cin = inputs.shape
    with store.variable_scope(scope):
        w = _weights((kh, kw, cin, num_output_channels), kh * kw * cin, kh * kw * num_output_channels, use_xavier,
                     stddev, weight_decay)
        bias = store.get_variable("biases", (num_output_channels,), constant_init(0.0))
        if kw == 1:
            x_, w_ = inputs, w.view(cin, num_output_channels)
        elif kw == wdt and padding == "VALID":
            x_, w_ = inputs.reshape(b_, h, 1, wdt * cin), w.reshape(kw * cin, num_output_channels)
        else:
            raise NotImplementedError(f"conv2d mirror: kernel [1,{kw}] on width {wdt} with padding {padding}")
        if bn:
            fused = _fused_dense_bn(x_, w_, bias, "bn", bn_decay, is_training, activation_fn)
            if fused is not None:
                return fused
        if not bn:
            return _dense_bias(x_, w_, bias, activation_fn)
        out = torch.matmul(x_, w_) + bias
        out = batch_norm_for_conv2d(out, is_training, bn_decay, "bn", data_format)
        if activation_fn is not None:
            out = activation_fn(out)
    return out


def fully_connected(inputs, num_outputs, scope, use_xavier=True, stddev=1e-3, weight_decay=None, activation_fn=relu,
                    bn=False, bn_decay=None, is_training=None):
    """inputs (B, N) (reference tf_util.py:318-355)."""
    store = get_default_store()
    cin = inputs.shape[-1]
    with store.variable_scope(scope):
        w = _weights((cin, num_outputs), cin, num_outputs, use_xavier, stddev, weight_decay)
        b = store.get_variable("biases", (num_outputs,), constant_init(0.0))
        if bn:
            fused = _fused_dense_bn(inputs, w, b, "bn", bn_decay, is_training, activation_fn)
            if fused is not None:
                return fused
        if not bn:
            return _dense_bias(inputs, w, b, activation_fn)
        out = torch.matmul(inputs, w) + b
        out = batch_norm_for_fc(out, is_training, bn_decay, "bn")
        if activation_fn is not None:
            out = activation_fn(out)
    return out


def dropout(inputs, is_training, scope, keep_prob=0.5, noise_shape=None):
    """tf.nn.dropout(inputs, keep_prob) when training (reference tf_util.py:590-615).  On CUDA the mask comes from the
    library's counter-based generator (seeded by the scope name and the store's seed, advanced by the optimizer step)."""
    if noise_shape is not None:
        raise NotImplementedError("noise_shape is not used on the hot path")
    if not bool(is_training):
        return inputs
    if inputs.is_cuda and inputs.dtype == torch.float32:
        import zlib
        from . import _ops
        store = get_default_store()
        seed = (zlib.crc32(store.full_name(scope).encode()) << 20) ^ int(getattr(store, "seed", 0))
        return _ops.dropout(inputs, keep_prob, seed)
    return F.dropout(inputs, p=1.0 - keep_prob, training=True)


def _dense_bias(x, w2d, bias, activation_fn):
    """x @ w + b (+ activation) for a layer without batch norm; the library's GEMMs on CUDA."""
    if x.is_cuda and x.dtype == torch.float32:
        from . import _ops
        out = _ops.DenseBias.apply(x, w2d, bias)
    else:
        out = torch.matmul(x, w2d) + bias
    return activation_fn(out) if activation_fn is not None else out
