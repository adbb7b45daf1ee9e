"""Small driver for compute-sanitizer: one call of every kernel added after the first sanitizer run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
torch.manual_seed(0)
dev = "cuda"
# persistent rows GEMM (NT = 32, 64, 128, 160), with and without fused statistics, ragged rows / k
for rows, k, n in [(38017, 70, 48), (38000, 128, 134), (40000, 9, 64), (37900, 134, 128), (38000, 3, 32)]:
    a = torch.randn(rows, k, device=dev); bt = torch.randn(n, k, device=dev)
    ws = torch.zeros(2 * n, device=dev, dtype=torch.float64)
    c = _ops.gemm_rows(a, bt); c2, _ = _ops.gemm_rows(a, bt, col_stats=ws)
    assert torch.allclose(c, a @ bt.t(), rtol=1e-3, atol=1e-3)
# one-tile-per-CTA rows GEMM
a = torch.randn(5000, 134, device=dev); bt = torch.randn(128, 134, device=dev); _ops.gemm_rows(a, bt)
# weight gradients: TMEM-fed (P = X and P = Y), short-reduction kernel, narrow kernel
for rows, m, n in [(40001, 134, 128), (39000, 64, 64), (38000, 128, 40), (16384, 256, 512), (8192, 64, 64), (40000, 3, 32),
                   (37000, 32, 3), (36000, 16, 16), (5000, 9, 130)]:
    x = torch.randn(rows, m, device=dev); y = torch.randn(rows, n, device=dev)
    dw = _ops.gemm_wgrad(x, y)
    assert torch.allclose(dw, x.t() @ y, rtol=1e-3, atol=2e-2)
# Non-Local forward / backward on tcgen05 (d = 32), ragged tails and the key-split path
for b, s, n in [(2, 200, 333), (1, 130, 4100)]:
    q = torch.randn(b, s, 32, device=dev, requires_grad=True); kv = torch.randn(b, n, 64, device=dev, requires_grad=True)
    out = _ops.NLCore.apply(q, kv, 32 ** -0.5)
    out.backward(torch.randn_like(out))
# PointConv contraction and max-pool
x = torch.randn(2, 40, 32, 70, device=dev, requires_grad=True); w = torch.randn(2, 40, 32, 32, device=dev, requires_grad=True)
_ops.pointconv_contract(x, w).sum().backward()
z = torch.randn(2, 33, 17, 20, device=dev, requires_grad=True)
_ops.MaxPoolK.apply(z).sum().backward()
torch.cuda.synchronize()
print("sanitize_new_kernels: done")
