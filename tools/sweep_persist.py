"""Time pasnl_gemm_rows_tf32x3 on the big-row shapes of the cls / seg steps (run twice: PASNL_GEMM_PERSIST=0/1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
shapes = [(262144, 128, 128), (524288, 64, 64), (262144, 134, 128), (262144, 128, 134), (524288, 9, 64), (524288, 3, 32),
          (262144, 32, 3), (65536, 128, 128), (40000, 64, 200)]
for rows, k, n in shapes:
    a = torch.randn(rows, k, device="cuda"); bt = torch.randn(n, k, device="cuda")
    ws = torch.zeros(2 * n, device="cuda", dtype=torch.float64)
    med, _ = timeit(lambda: _ops.gemm_rows(a, bt), iters=10)
    meds, _ = timeit(lambda: _ops.gemm_rows(a, bt, col_stats=ws), iters=10)
    want = a.double() @ bt.double().t()
    got = _ops.gemm_rows(a, bt)
    err = float((got.double() - want).abs().max()) / float(want.abs().max())
    gb = 4.0 * rows * (k + n) / 1e9
    print(f"rows={rows:7d} k={k:4d} n={n:4d}  {med:7.1f} us ({gb / med * 1e6:6.0f} GB/s)  with stats {meds:7.1f} us  relerr {err:.1e}", flush=True)
