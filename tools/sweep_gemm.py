import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
rows = 262144
for n in (64, 128):
    for k in (16, 32, 64, 128, 256, 512):
        a = torch.randn(rows, k, device="cuda"); bt = torch.randn(n, k, device="cuda")
        med, best = timeit(lambda: _ops.gemm_rows(a, bt), iters=10)
        x = torch.randn(rows, k, device="cuda"); dy = torch.randn(rows, n, device="cuda")
        medw, _ = timeit(lambda: _ops.gemm_wgrad(x, dy), iters=10) if k >= 32 else (0, 0)
        medc, _ = timeit(lambda: torch.mm(a, bt.t()), iters=10)
        print(f"rows={rows} k={k:4d} n={n:4d}  rows-gemm {med:7.1f} us  wgrad {medw:7.1f} us  cublas-fp32 {medc:7.1f} us  bytes {4*rows*(k+n)/1e6:6.1f} MB -> {4*rows*(k+n)/med/1e6:6.2f} TB/s")
