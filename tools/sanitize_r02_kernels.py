"""compute-sanitizer driver for the kernels added in round 2 (one call each, ragged shapes):
   compute-sanitizer --tool memcheck python tools/sanitize_r02_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointasnl_b200 import _lib, _ops, tf_grouping, tf_sampling
from pointasnl_b200.grid_subsampling import grid_subsampling

torch.manual_seed(0)
dev = "cuda"
# weights-stationary TMA rows GEMM (both weight layouts, prologue, ragged rows / k / n) and the mid-size TMA kernel
for rows, k, n, w_kn in [(37999, 100, 70, True), (40001, 128, 128, False), (38017, 36, 72, True)]:
    a = torch.randn(rows, k, device=dev)
    w = torch.randn(k, n, device=dev) if w_kn else torch.randn(n, k, device=dev)
    st = torch.zeros(2 * n, dtype=torch.float64, device=dev)
    c, _ = _ops.gemm_rows(a, w, w_kn=w_kn, col_stats=st)
    want = a @ (w if w_kn else w.t())
    assert torch.allclose(c, want, rtol=1e-3, atol=1e-3)
a = torch.randn(45001, 96, device=dev)
w = torch.randn(96, 72, device=dev)
c = torch.empty(45001, 80, device=dev)
sc, sh = torch.rand(96, device=dev) + 0.5, torch.randn(96, device=dev)
_lib.call("pasnl_gemm_rows_ws", 45001, 96, 72, a.data_ptr(), 96, w.data_ptr(), 72, 1, c.data_ptr(), 80, None, sc.data_ptr(),
          sh.data_ptr(), 1, _ops._stream())
for rows, k, n in [(12700, 516, 130), (3000, 1100, 1024)]:
    a = torch.randn(rows, k, device=dev)
    bt = torch.randn(n, k, device=dev)
    assert torch.allclose(_ops.gemm_rows(a, bt), a @ bt.t(), rtol=1e-3, atol=2e-2)
# one-tile kernel with split-k and narrowed tiles; 134-column dgrad split
a = torch.randn(4097, 2500, device=dev)
bt = torch.randn(70, 2500, device=dev)
assert torch.allclose(_ops.gemm_rows(a, bt), a @ bt.t(), rtol=1e-3, atol=5e-2)
dy = torch.randn(38000, 128, device=dev)
w = torch.randn(134, 128, device=dev)
assert torch.allclose(_ops.gemm_rows(dy, w), dy @ w.t(), rtol=1e-3, atol=1e-2)
# MN-major TMA weight gradient (ragged rows, partial m / n tiles)
for rows, m, n in [(5001, 100, 72), (70001, 132, 96), (9999, 260, 40), (2049, 32, 32)]:
    x = torch.randn(rows, m, device=dev)
    y = torch.randn(rows, n, device=dev)
    assert torch.allclose(_ops.gemm_wgrad(x, y), x.t() @ y, rtol=1e-3, atol=5e-2)
# small ops
x = torch.randn(33, 777, device=dev, requires_grad=True)
w = torch.randn(777, 130, device=dev, requires_grad=True)
b = torch.randn(130, device=dev, requires_grad=True)
_ops.DenseBias.apply(x, w, b).sum().backward()
p, q, r = [torch.randn(5, 7, 13, device=dev, requires_grad=True) for _ in range(3)]
(_ops.add_n(p, q, r).sum() + _ops.concat_cols(p, q).sum()).backward()
_ops.split_cols(torch.randn(9, 11, 6, device=dev), 3)
step = torch.zeros((), dtype=torch.int64, device=dev)
_ops.set_dropout_step(step)
d = torch.randn(1001, device=dev, requires_grad=True)
_ops.dropout(d, 0.4, 7).sum().backward()
_ops.set_dropout_step(None)
lg = torch.randn(4, 333, 21, device=dev, requires_grad=True)
_ops.softmax_ce(lg, torch.randint(0, 21, (4, 333), device=dev), torch.rand(4, 333, device=dev), 2.0).backward()
g = torch.randn(2, 37, 16, 11, device=dev, requires_grad=True)
pooled, ident, xyz = _ops.SAFront.apply(g)
(pooled.sum() + (ident * 2).sum() + xyz.sum()).backward()
n = 10007
prm, grad, m1, v1 = [torch.randn(n, device=dev) for _ in range(4)]
v1.abs_()
_ops.adam_l2(prm, grad, m1, v1, (torch.rand(n, device=dev) < 0.5).to(torch.uint8), 1e-4, 1.0, 1e-3, step,
             torch.zeros((), dtype=torch.float64, device=dev), torch.zeros((), device=dev), torch.zeros((), device=dev))
tr = _ops.WeightTransposer()
tr.register([torch.randn(1, 1, 134, 128, device=dev), torch.randn(259, 7, device=dev)])
tr.refresh()
# a12 ops and grid subsampling
tf_sampling.prob_sample(torch.rand(3, 8193, device=dev), torch.rand(3, 100, device=dev))
tf_grouping.select_top_k(5, torch.randn(2, 9, 33, device=dev))
pts = (np.random.default_rng(0).normal(size=(30001, 3)) * np.array([2.0, 1.5, 0.5])).astype(np.float32)
grid_subsampling(torch.from_numpy(pts).to(dev), features=torch.rand(30001, 3, device=dev),
                 labels=torch.randint(0, 9, (30001,), device=dev, dtype=torch.int32), sampleDl=0.1)
torch.cuda.synchronize()
print("sanitize_r02_kernels: done")
