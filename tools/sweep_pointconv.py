import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
for g, k, c in [(16384, 32, 64), (4096, 64, 128), (40960, 32, 16)]:
    x = torch.randn(g, k, c, device="cuda", requires_grad=True); w = torch.randn(g, k, 32, device="cuda", requires_grad=True)
    with torch.no_grad():
        med, _ = timeit(lambda: _ops.pointconv_contract(x, w), iters=10)
    out = _ops.pointconv_contract(x, w); go = torch.randn_like(out)
    medb, _ = timeit(lambda: torch.autograd.grad(out, (x, w), go, retain_graph=True), iters=10)
    fb = 4.0 * g * (k * c + k * 32 + c * 32); bb = 4.0 * g * (2 * k * c + 2 * k * 32 + c * 32)
    print(f"pointconv G={g} K={k} C={c}: fwd {med:.1f} us ({fb / med / 1e3:.0f} GB/s)  bwd {medb:.1f} us ({bb / medb / 1e3:.0f} GB/s)", flush=True)
