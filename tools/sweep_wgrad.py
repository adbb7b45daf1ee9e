"""Time pasnl_gemm_wgrad_tf32x3 on the long-reduction shapes of the cls step (PASNL_GEMM_PERSIST=0: previous kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
shapes = [(1310720, 16, 16), (1310720, 9, 16), (262144, 128, 128), (262144, 134, 128), (524288, 64, 64), (16384, 2048, 128), (16384, 256, 512), (524288, 9, 64),
          (524288, 3, 32), (262144, 32, 3), (262144, 3, 32), (32768, 3, 64), (16384, 9, 128), (16384, 6, 32)]
for rows, m, n in shapes:
    x = torch.randn(rows, m, device="cuda"); y = torch.randn(rows, n, device="cuda")
    med, _ = timeit(lambda: _ops.gemm_wgrad(x, y), iters=10)
    medc, _ = timeit(lambda: torch.mm(x.t(), y), iters=10)
    want = x.double().t() @ y.double()
    got = _ops.gemm_wgrad(x, y)
    err = float((got.double() - want).abs().max()) / float(want.abs().max())
    gb = 4.0 * rows * (m + n) / 1e9
    print(f"rows={rows:7d} m={m:4d} n={n:4d}  {med:7.1f} us ({gb / med * 1e6:6.0f} GB/s)  cuBLAS {medc:7.1f} us  relerr {err:.1e}", flush=True)
