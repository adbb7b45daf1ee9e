import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
def tc(a, bt):
    rows, k = a.shape; n = bt.shape[0]
    c = torch.empty((rows, n), device=a.device)
    _ops._call("pasnl_gemm_rows_tf32x3", rows, k, n, _ops._ptr(a), k, _ops._ptr(bt), k, _ops._ptr(c), n, None, _ops._stream())
    return c
for rows, k, n in [(4096, 259, 256), (4096, 256, 259), (4096, 256, 256), (4096, 134, 256), (4096, 256, 134), (4096, 64, 256), (4096, 256, 64),
                   (4096, 131, 64), (4096, 1024, 256), (4096, 256, 1024), (4096, 512, 256), (4096, 4096, 256), (8192, 256, 256), (2048, 512, 512),
                   (1024, 1024, 512), (16384, 128, 128), (4096, 2048, 256)]:
    a = torch.randn(rows, k, device="cuda"); bt = torch.randn(n, k, device="cuda")
    if k > 1024:
        t1, _ = timeit(lambda: _ops.gemm_rows(a, bt), iters=8)   # python path (k slabs)
    else:
        t1, _ = timeit(lambda: tc(a, bt), iters=8)
    t2, _ = timeit(lambda: torch.mm(a, bt.t()), iters=8)
    print(f"rows={rows:6d} k={k:5d} n={n:5d}  tc {t1:7.1f} us  cuBLAS {t2:7.1f} us", flush=True)
