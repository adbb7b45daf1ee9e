import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
rows, m, n = [int(x) for x in sys.argv[1:4]]
x = torch.randn(rows, m, device="cuda"); y = torch.randn(rows, n, device="cuda")
for _ in range(3):
    c = _ops.gemm_wgrad(x, y)
torch.cuda.synchronize()
