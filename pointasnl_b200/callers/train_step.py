"""One training step of a caller model (forward + loss + backward + gradient all-reduce + Adam), eager or captured
into a CUDA graph.  This is the call bench.py times and what a user of the drop-in API would write by hand."""
import torch

from .. import tf_util
from ..dp import FlatGradBucket
from . import pointasnl_cls, pointasnl_sem_seg_res


class TrainStep:
    def __init__(self, model="cls", batch_size=32, num_point=1024, device="cuda", adaptive_sample=False,
                 num_class=None, feature_channel=3, lr=1e-3, seed=0, use_graph=False, radius=0.07, overlap_allreduce=False):
        """model: 'cls' (models/pointasnl_cls.py) or 'seg' (models/pointasnl_sem_seg_res.py; feature_channel=3, radius=0.07
        is the ScanNet configuration, feature_channel=1, num_class=20, radius=0.2 the SemanticKITTI one,
        SemanticKITTI/train_semantic_kitti.py:81, train_semantic_kitti_grid.py:144)."""
        self.kind = model
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("TrainStep drives the CUDA library (optimizer kernel included): it needs a CUDA device")
        self.batch_size, self.num_point = batch_size, num_point
        self.adaptive_sample = adaptive_sample
        self.feature_channel = feature_channel
        self.radius = radius
        self.num_class = num_class or (40 if model == "cls" else 21)
        self.store = tf_util.VariableStore(self.device, seed=seed)
        c = 6 if model == "cls" else 3 + feature_channel
        # static device buffers: the step reads these (required for graph replay, harmless in eager mode)
        self.points = torch.zeros((batch_size, num_point, c), device=self.device)
        if model == "cls":
            self.labels = torch.zeros((batch_size,), dtype=torch.int64, device=self.device)
            self.smpw = None
        else:
            self.labels = torch.zeros((batch_size, num_point), dtype=torch.int64, device=self.device)
            self.smpw = torch.ones((batch_size, num_point), device=self.device)
        self.loss = torch.zeros((), device=self.device)
        # materialise every variable (TF does this at graph-construction time) with an inference-mode pass: it creates
        # the same variables but neither updates the moving averages nor draws dropout masks
        with torch.no_grad():
            self._forward(is_training=False)
        self.bucket = FlatGradBucket(self.store.parameters())
        # ONE flat parameter buffer (all variables are views of it; the gradient bucket mirrors its layout) updated by ONE
        # library kernel per step: L2 regulariser on the 'weights' variables + TF-style Adam (pasnl_adam_l2).
        self.flat_param = self.bucket.flatten_parameters()
        self.lr = float(lr)
        self.l2_coeff = 1e-4                       # get_loss(weights_decay=1e-4) of both reference models
        dev = self.device
        self.adam_m = torch.zeros_like(self.flat_param.data)
        self.adam_v = torch.zeros_like(self.flat_param.data)
        self.adam_step = torch.zeros((), dtype=torch.int64, device=dev)
        self.reg_loss = torch.zeros((), dtype=torch.float64, device=dev)
        self.l2_mask = torch.zeros(self.flat_param.numel(), dtype=torch.uint8, device=dev)
        by_param = {id(v): name for name, v in self.store.variables().items()}
        off = 0
        weight_params = []
        for prm in self.bucket.params:
            name = by_param.get(id(prm), "")
            if "weights" in name:                  # tf.global_variables() with 'weights' in the name (pointasnl_cls.py:58)
                self.l2_mask[off:off + prm.numel()] = 1
                if prm.dim() >= 2:
                    weight_params.append(prm)
            off += prm.numel()
        from .. import _ops
        # zero-initialised scratch of the dense glue, cleared once per step instead of ~100 separate fill nodes
        self.arena = _ops.ZeroArena(self.device) if self.device.type == "cuda" else None
        # K-major copies of every weight matrix, refreshed by one batched launch at the start of a step
        self.transposer = None
        if self.device.type == "cuda":
            self.transposer = _ops.WeightTransposer()
            self.transposer.register(weight_params)
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = None
        # data parallel, optional (overlap_allreduce=True or PASNL_DP_OVERLAP=1): the gradient bucket is all-reduced in two
        # pieces.  Backward visits the layers in reverse creation order, so the gradients of the LATER ~65 % of the bucket
        # are complete long before the first layers' are: a multi-grad hook ships that piece (async NCCL all-reduce,
        # captured into the step's single CUDA graph) while the early layers' backward is still running.
        # OFF by default -- measured on 2 x B200 (cls, B=32/GPU): 5.98 ms/step with the overlap vs 5.86 ms with one
        # all-reduce after the backward (1-GPU step 5.82): NCCL's CTAs take SMs from the persistent one-CTA-per-SM GEMMs
        # they run beside, which costs more than the 13 MB exchange itself.
        import os
        env = os.environ.get("PASNL_DP_OVERLAP")
        self.overlap_allreduce = bool(overlap_allreduce) if env is None else env[:1] == "1"
        total, acc, self.split = self.bucket.numel, 0, 0
        for i, prm in enumerate(self.bucket.params):
            if acc >= 0.35 * total:
                self.split = i
                break
            acc += prm.numel()
        self.split_off = self.bucket.offset_of(self.split)
        self._works = []

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _ship_late_gradients(self, grads):
        import torch.distributed as dist
        self.bucket.pack_grads(self.split, grads)
        self._works.append(dist.all_reduce(self.bucket.flat[self.split_off:], op=dist.ReduceOp.SUM, async_op=True))

    def _forward(self, is_training=True, weights_decay=1e-4):
        with self.store.as_default():
            if self.kind == "cls":
                pred, end_points = pointasnl_cls.get_model(self.points, is_training, use_normal=True,
                                                           adaptive_sample=self.adaptive_sample,
                                                           num_class=self.num_class)
                return pointasnl_cls.get_loss(pred, self.labels, end_points, weights_decay=weights_decay)
            pred, end_points = pointasnl_sem_seg_res.get_model(self.points, is_training, self.num_class,
                                                               feature_channel=self.feature_channel)
            return pointasnl_sem_seg_res.get_loss(pred, self.labels, end_points, self.smpw, radius=self.radius,
                                                  weights_decay=weights_decay)

    def _fwd_bwd(self):
        from .. import _ops
        self.bucket.detach()
        if self.arena is not None:
            self.arena.reset()
        prev_t, prev_d = _ops._transposer, _ops._dropout_step
        _ops.set_weight_transposer(self.transposer)
        _ops.set_dropout_step(self.adam_step)
        try:
            if self.transposer is not None:
                self.transposer.refresh()
            overlap = self.overlap_allreduce and self._world() > 1 and 0 < self.split < len(self.bucket.params)
            handle = None
            with _ops.use_zero_arena(self.arena):
                loss = self._forward(weights_decay=0.0)      # the regulariser is folded into the optimizer kernel
                if overlap:
                    handle = torch.autograd.graph.register_multi_grad_hook(tuple(self.bucket.params[self.split:]),
                                                                           self._ship_late_gradients)
                loss.backward()
            if handle is not None:
                handle.remove()
        finally:
            _ops.set_weight_transposer(prev_t)
            _ops.set_dropout_step(prev_d)
        if overlap and self._works:
            import torch.distributed as dist
            self.bucket.pack(0, self.split)
            self._works.append(dist.all_reduce(self.bucket.flat[:self.split_off], op=dist.ReduceOp.SUM, async_op=True))
            for w in self._works:
                w.wait()
            self._works = []
            self._reduced = True
        else:
            self.bucket.pack()
            self._reduced = False
        self.data_loss = loss.detach()

    def _optimizer_step(self):
        """L2 + Adam over the flat buffers; also publishes loss = data loss + regulariser and bumps the step counter."""
        import torch.distributed as dist
        from .. import _ops
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        _ops.adam_l2(self.flat_param.data, self.bucket.flat, self.adam_m, self.adam_v, self.l2_mask, self.l2_coeff,
                     1.0 / world, self.lr, self.adam_step, self.reg_loss, self.data_loss, self.loss)

    def _eager_step(self):
        self._fwd_bwd()
        if not self._reduced:
            self.bucket.all_reduce_sum()
        self._optimizer_step()

    def capture(self, warmup=3):
        """Capture the step.  Single GPU / no overlap: fwd+bwd and the optimizer as two graphs around the (uncaptured)
        NCCL all-reduce.  Data parallel with overlap: ONE graph holding fwd + bwd + both all-reduce pieces (NCCL is
        captured on its own stream, forked from the backward) + the optimizer kernel; if NCCL capture is not possible
        the two-graph form without overlap is used."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from .. import _lib
        c0 = _lib.launch_count
        self.g_all = None
        if self.overlap_allreduce and self._world() > 1:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._fwd_bwd()
                    self._optimizer_step()
                self.g_all = g
                self.launches_per_step = _lib.launch_count - c0
                self.graph = True
                return
            except Exception as e:       # NCCL not capturable in this environment: fall back, loudly
                import sys
                print(f"[pasnl] all-reduce overlap disabled (graph capture of NCCL failed: {e!r})", file=sys.stderr)
                self.overlap_allreduce = False
                self._works = []
                torch.cuda.synchronize()
                c0 = _lib.launch_count
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self._fwd_bwd()
        self.launches_per_step = _lib.launch_count - c0   # pasnl kernels captured in (and replayed by) one step
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt):
            self._optimizer_step()
        self.graph = True

    def set_inputs(self, points, labels, smpw=None, non_blocking=True):
        """Copy one batch (host pinned or device) into the static buffers."""
        self.points.copy_(points, non_blocking=non_blocking)
        self.labels.copy_(labels, non_blocking=non_blocking)
        if smpw is not None and self.smpw is not None:
            self.smpw.copy_(smpw, non_blocking=non_blocking)

    def step(self):
        if self.graph and getattr(self, "g_all", None) is not None:
            self.g_all.replay()
        elif self.graph:
            self.g_fb.replay()
            self.bucket.all_reduce_sum()
            self.g_opt.replay()
        else:
            self._eager_step()
        return self.loss
