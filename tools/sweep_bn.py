import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointasnl_b200 import _ops
from tools.bench_op import timeit
for rows, c in [(262144, 128), (524288, 64), (524288, 32), (16384, 128), (16384, 512), (4096, 256), (4096, 1024)]:
    x = torch.randn(rows, 64, device="cuda", requires_grad=True)
    w = torch.randn(64, c, device="cuda", requires_grad=True)
    b = torch.zeros(c, device="cuda", requires_grad=True); g = torch.ones(c, device="cuda", requires_grad=True); be = torch.zeros(c, device="cuda", requires_grad=True)
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    out = _ops.DenseBNAct.apply(x, w, b, g, be, rm, rv, 0.9, 1e-3, 1)
    y, mean, var = out.grad_fn.saved_tensors[2], out.grad_fn.saved_tensors[5], out.grad_fn.saved_tensors[6]
    dout = torch.randn_like(out)
    dy = torch.empty_like(y); dg = torch.empty_like(g); db = torch.empty_like(be)
    ws = torch.empty(2 * c, device="cuda", dtype=torch.float64)
    def f():
        _ops._call("pasnl_bn_act_bwd", rows, c, _ops._ptr(y), _ops._ptr(dout), _ops._ptr(g), _ops._ptr(be), _ops._ptr(mean), _ops._ptr(var), 1e-3, 1,
                   _ops._ptr(dy), _ops._ptr(dg), _ops._ptr(db), 0, _ops._ptr(ws), _ops._stream())
    med, best = timeit(f, iters=10)
    print(f"bn_act_bwd rows={rows} c={c}: {med:.1f} us ({20.0 * rows * c / med / 1e3:.0f} GB/s)", flush=True)
