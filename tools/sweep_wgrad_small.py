import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
def tc(x, y):
    rows, m = x.shape; n = y.shape[1]
    dw = torch.empty((m, n), device=x.device)
    _ops._call("pasnl_gemm_wgrad_tf32x3", rows, m, n, _ops._ptr(x), m, _ops._ptr(y), n, _ops._ptr(dw), n, 0, _ops._stream())
    return dw
for rows, m, n in [(4096, 1024, 512), (4096, 512, 1024), (4096, 4096, 256), (4096, 2048, 256), (4096, 256, 4096), (4096, 1024, 256), (8192, 512, 512),
                   (4096, 259, 256), (4096, 256, 512), (2048, 512, 512), (1024, 1024, 512)]:
    x = torch.randn(rows, m, device="cuda"); y = torch.randn(rows, n, device="cuda")
    a, _ = timeit(lambda: tc(x, y), iters=8)
    b, _ = timeit(lambda: torch.mm(x.t(), y), iters=8)
    err = float((tc(x, y) - x.t() @ y).abs().max())
    print(f"rows={rows:6d} m={m:5d} n={n:5d}  tc {a:7.1f} us  cuBLAS {b:7.1f} us   maxdiff {err:.2e}", flush=True)
