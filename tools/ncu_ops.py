"""Launches ONE library entry point a few times (for `ncu -k regex:... --launch-skip N -c 1`).  python tools/ncu_ops.py <op>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pointasnl_b200 import _ops  # noqa: E402

# op -> (entry point, dims, kernel-name regex, launcher)
OPS = {}


def op(name, entry, dims, regex):
    def deco(fn):
        OPS[name] = (entry, dims, regex, fn)
        return fn
    return deco


@op("rows_ws_262144x128x128", "pasnl_gemm_rows_ws", [262144, 128, 128], "gemm_rows_ws_kernel")
def _rows_ws():
    a = torch.randn(262144, 128, device="cuda")
    w = torch.randn(128, 128, device="cuda") / 11
    st = torch.zeros(256, dtype=torch.float64, device="cuda")
    return lambda: _ops.gemm_rows(a, w, w_kn=True, col_stats=st)


@op("rows_ws_524288x64x64", "pasnl_gemm_rows_ws", [524288, 64, 64], "gemm_rows_ws_kernel")
def _rows_ws2():
    a = torch.randn(524288, 64, device="cuda")
    w = torch.randn(64, 64, device="cuda") / 8
    st = torch.zeros(128, dtype=torch.float64, device="cuda")
    return lambda: _ops.gemm_rows(a, w, w_kn=True, col_stats=st)


@op("wgrad_262144x128x128", "pasnl_gemm_wgrad_tf32x3", [262144, 128, 128], "gemm_wgrad_tma_kernel")
def _wgrad():
    x = torch.randn(262144, 128, device="cuda")
    dy = torch.randn(262144, 128, device="cuda")
    return lambda: _ops.gemm_wgrad(x, dy)


@op("rows_16384x1024x128", "pasnl_gemm_rows_tf32x3", [16384, 1024, 128], "gemm_rows_tma_kernel")
def _rows_mid():
    a = torch.randn(16384, 1024, device="cuda")
    bt = torch.randn(128, 1024, device="cuda") / 32
    return lambda: _ops.gemm_rows(a, bt)


@op("nl_fwd_cls_l1", "pasnl_nl_core_fwd", [32, 512, 1024, 32], "nl_fwd_tc_kernel")
def _nl_fwd():
    q = torch.randn(32, 512, 32, device="cuda")
    kv = torch.randn(32, 1024, 64, device="cuda")
    return lambda: _ops.NLCore.apply(q, kv, 32 ** -0.5)


@op("nl_bwd_dq_cls_l1", "pasnl_nl_core_bwd", [32, 512, 1024, 32], "nl_bwd_dq_tc_kernel")
def _nl_bwd_dq():
    q = torch.randn(32, 512, 32, device="cuda", requires_grad=True)
    kv = torch.randn(32, 1024, 64, device="cuda", requires_grad=True)
    out = _ops.NLCore.apply(q, kv, 32 ** -0.5)
    g = torch.randn_like(out)
    return lambda: torch.autograd.grad(out, (q, kv), g, retain_graph=True)


OPS["nl_bwd_dkv_cls_l1"] = ("pasnl_nl_core_bwd", [32, 512, 1024, 32], "nl_bwd_dkv_tc_kernel", _nl_bwd_dq)


@op("bn_bwd_262144x128", "pasnl_bn_act_bwd", [262144, 128], "bn_bwd_apply_kernel")
def _bn_bwd():
    x = torch.randn(262144, 16, device="cuda", requires_grad=True)
    w = (torch.randn(16, 128, device="cuda") / 4).requires_grad_(True)
    z = [torch.zeros(128, device="cuda", requires_grad=True), torch.ones(128, device="cuda", requires_grad=True),
         torch.zeros(128, device="cuda", requires_grad=True)]
    y = _ops.DenseBNAct.apply(x, w, z[0], z[1], z[2], torch.zeros(128, device="cuda"), torch.ones(128, device="cuda"), 0.9, 1e-3, 1)
    g = torch.randn_like(y)
    return lambda: torch.autograd.grad(y, (x, w), g, retain_graph=True)


if __name__ == "__main__":
    fn = OPS[sys.argv[1]][3]()
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
