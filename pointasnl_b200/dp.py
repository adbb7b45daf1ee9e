"""Data parallelism for the callers: one process per GPU, batches sharded by rank, ONE all-reduce per step.

The reference is single-GPU only (no collective call sites), so this is new functionality; its
contract is "N-rank gradients == the mean of the per-rank gradients" (per-replica batch-norm statistics, i.e. plain
data-parallel replicas).  Every hot-path kernel is per-cloud, so there is no collective inside the path: the only
exchange is the gradient bucket at the optimizer boundary -- all parameter gradients live in ONE flat fp32 buffer
(``p.grad`` are views into it), reduced with a single NCCL all-reduce over NVLink/NVSwitch (gloo in CPU tests).
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """Owns one contiguous fp32 buffer holding every parameter's gradient; ``p.grad`` alias slices of it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("all parameters must be float32 on one device")
            off += p.numel()
        self.views = list(torch.split(self.flat, [p.numel() for p in self.params]))
        self.views = [v.view_as(p) for v, p in zip(self.views, self.params)]
        self.alias()

    def alias(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()

    def detach(self):
        """Drop ``p.grad`` before backward so autograd ASSIGNS each gradient instead of accumulating into the bucket
        (accumulation is one zero-fill plus one add kernel per parameter per step); follow with ``pack()``."""
        for p in self.params:
            p.grad = None

    def pack(self, start=0, stop=None):
        """Copy the freshly produced gradients of params[start:stop] into the bucket with one multi-tensor copy and
        re-alias every ``p.grad``."""
        sel = list(zip(self.views, self.params))[start:stop]
        have = [(v, p.grad) for v, p in sel if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in sel:
            if p.grad is None:
                v.zero_()     # parameter not reached by this step's loss
        self.alias()

    def pack_grads(self, start, grads):
        """Copy `grads` (as handed to an autograd multi-grad hook: one per params[start:], None = not reached) into the
        bucket.  Used to ship the late layers' gradients while the early layers' backward is still running."""
        views = self.views[start:start + len(grads)]
        have = [(v, g) for v, g in zip(views, grads) if g is not None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, g in zip(views, grads):
            if g is None:
                v.zero_()

    def offset_of(self, index):
        return sum(p.numel() for p in self.params[:index])

    def check_aliasing(self):
        """autograd must have accumulated in place (it does when .grad is pre-populated); guard against a silent
        re-allocation that would detach a gradient from the bucket."""
        base = self.flat.data_ptr()
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                raise RuntimeError("a parameter gradient no longer aliases the flat bucket")
            off += p.numel()

    def all_reduce_mean(self, group=None, async_op=False):
        """Single collective: sum over ranks, then divide by the world size."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(group)
        if world == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work, world
        self.flat.div_(world)
        return None

    def all_reduce_sum(self, group=None):
        """Single collective, sum only: the 1/world factor is folded into the optimizer kernel (pasnl_adam_l2's
        grad_scale), which saves the separate division pass over the bucket."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return None

    def nbytes(self):
        return self.numel * 4

    def flatten_parameters(self):
        """Move every parameter's storage into ONE flat fp32 buffer (``p.data`` become views of it, values kept) and
        return that buffer wrapped as a single Parameter whose ``.grad`` is the gradient bucket: the optimizer then
        updates 1 tensor with 1 kernel instead of one multi-tensor launch chain over ~100 small tensors."""
        flat = torch.empty(self.numel, device=self.flat.device, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view_as(p)
                off += n
        fp = torch.nn.Parameter(flat, requires_grad=True)
        fp.grad = self.flat
        return fp


def shard_range(total, rank, world):
    """Rank r's slice [lo, hi) of `total` independent clouds (strong scaling); remainders go to the low ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device):
    """MAX-reduce a python float over ranks (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
