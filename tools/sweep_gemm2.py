import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
shapes = [(4096,256,512),(4096,512,1024),(4096,1024,512),(4096,1024,256),(4096,4096,256),(16384,2048,128),(16384,128,256),(16384,256,512),(16384,512,256),(4096,259,256),(4096,256,256),(16384,131,128),(16384,128,128),(4096,512,256),(262144,134,128),(524288,9,64),(524288,64,64),(524288,3,32),(262144,3,32),(32768,3,64),(16384,6,32),(16384,32,128),(16384,9,128)]
for rows,k,n in shapes:
    a = torch.randn(rows, k, device="cuda"); bt = torch.randn(n, k, device="cuda")
    _ops.GEMM_IMPL = "tc"
    med, _ = timeit(lambda: _ops.gemm_rows(a, bt), iters=8)
    medc, _ = timeit(lambda: torch.mm(a, bt.t()), iters=8)
    x = torch.randn(rows, k, device="cuda"); dy = torch.randn(rows, n, device="cuda")
    medw, _ = timeit(lambda: _ops.gemm_wgrad(x, dy), iters=8)
    medwc, _ = timeit(lambda: torch.mm(x.t(), dy), iters=8)
    print(f"rows={rows:7d} k={k:5d} n={n:5d}  fwd tc {med:7.1f} cublas {medc:7.1f} | wgrad tc {medw:7.1f} cublas {medwc:7.1f}")
