"""Mid-size layers (rows 4096..16384) on the one-tile-per-CTA rows GEMM vs cuBLAS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
shapes = [(16384, 128, 2048), (16384, 1024, 128), (4096, 256, 4096), (4096, 1024, 512), (4096, 512, 1024), (16384, 256, 512),
          (16384, 512, 256), (16384, 128, 128), (16384, 128, 256)]
for rows, k, n in shapes:
    a = torch.randn(rows, k, device="cuda"); bt = torch.randn(n, k, device="cuda")
    med, _ = timeit(lambda: _ops.gemm_rows(a, bt), iters=10)
    medc, _ = timeit(lambda: torch.mm(a, bt.t()), iters=10)
    print(f"rows={rows:6d} k={k:5d} n={n:5d}  tc {med:7.1f} us ({2.0 * rows * k * n / med / 1e6:5.1f} TF)  cuBLAS {medc:7.1f} us", flush=True)
