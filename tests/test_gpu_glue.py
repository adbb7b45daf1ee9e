"""Fused dense-glue epilogue (csrc/dense_glue.cu) against the op-by-op PyTorch formulation of the reference's
conv -> bias_add -> batch_norm(train, eps 1e-3, biased variance) -> activation (utils/tf_util.py:120-185,512-531)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_layer(x, w, b, gamma, beta, act):
    y = x.reshape(-1, w.shape[0]) @ w + b
    mean, var = y.mean(0), y.var(0, unbiased=False)
    z = (y - mean) / torch.sqrt(var + 1e-3) * gamma + beta
    if act == 1:
        z = torch.relu(z)
    elif act == 2:
        z = torch.nn.functional.leaky_relu(z, 0.2)
    return z.reshape(x.shape[:-1] + (w.shape[1],)), mean, var


@pytest.mark.parametrize("shape,cout,act", [((4, 128, 64, 134), 128, 1), ((8, 512, 32, 9), 64, 1), ((32, 1536), 512, 1),
                                            ((5, 128, 64, 134), 128, 0), ((3, 511, 32, 64), 70, 2),   # (act 0: no ReLU-kink flips among 5M outputs)
                                            ((3, 50, 7, 35), 66, 0), ((2, 777, 5), 3, 2), ((6, 40, 1, 2048), 128, 1)])
def test_fused_dense_bn_act_forward_backward(cuda, shape, cout, act):
    from pointasnl_b200 import _ops
    torch.manual_seed(cout)
    cin = shape[-1]
    x = torch.randn(*shape) * 1.7 + 0.3
    w = torch.randn(cin, cout) / cin ** 0.5
    b = torch.randn(cout)
    gamma, beta = torch.rand(cout) + 0.5, torch.randn(cout) * 0.2
    rm, rv = torch.zeros(cout), torch.ones(cout)
    xd, wd, bd, gd, bed = [t.double().requires_grad_(True) for t in (x, w, b, gamma, beta)]
    want, mean, var = reference_layer(xd, wd, bd, gd, bed, act)
    xc, wc, bc, gc, bec = [t.to(cuda).requires_grad_(True) for t in (x, w, b, gamma, beta)]
    rmc, rvc = rm.to(cuda), rv.to(cuda)
    got = _ops.DenseBNAct.apply(xc, wc, bc, gc, bec, rmc, rvc, 0.9, 1e-3, act)
    torch.testing.assert_close(got.detach().cpu().double(), want.detach(), rtol=1e-4, atol=5e-5)
    # moving statistics as TF keeps them: decay*old + (1-decay)*batch, mean includes the conv bias
    torch.testing.assert_close(rmc.cpu().double(), 0.1 * mean.detach(), rtol=1e-4, atol=1e-5)
    # (DenseBNAct's default = TF's fused batch-norm kernel: the moving variance is fed the Bessel-corrected variance)
    nrows = x.numel() // cin
    torch.testing.assert_close(rvc.cpu().double(), 0.9 + 0.1 * var.detach() * nrows / (nrows - 1), rtol=1e-4, atol=1e-5)
    g = torch.randn(got.shape)
    want.backward(g.double())
    got.backward(g.to(cuda))
    for name, a, r in (("x", xc, xd), ("w", wc, wd), ("gamma", gc, gd), ("beta", bec, bed)):
        err = (a.grad.cpu().double() - r.grad).norm() / r.grad.norm().clamp(min=1e-12)
        assert err < 1e-4, (name, float(err))
    assert float(bc.grad.abs().max()) == 0.0          # exact: the bias cancels inside train-mode batch norm
    assert float(bd.grad.abs().max()) < 1e-9 * float(wd.grad.abs().max()) + 1e-12


def test_tf_util_layers_take_the_fused_path_and_match_op_by_op(cuda):
    from pointasnl_b200 import _lib, tf_util
    torch.manual_seed(0)
    x = torch.randn(16, 64, 16, 20, device=cuda)       # 16384 rows: the size class _ops.gemm_rows sends to tcgen05
    store = tf_util.VariableStore(cuda, seed=5)
    with store.as_default():
        c0 = _lib.launch_count
        fused = tf_util.conv2d(x, 24, [1, 1], "c", padding="VALID", bn=True, is_training=True)
        assert _lib.launch_count - c0 <= 3            # [weight transpose +] tcgen05 GEMM with fused statistics + normalise/activate
        orig = tf_util._fused_dense_bn
        tf_util._fused_dense_bn = lambda *a, **k: None
        try:
            plain = tf_util.conv2d(x, 24, [1, 1], "c", padding="VALID", bn=True, is_training=True)
        finally:
            tf_util._fused_dense_bn = orig
    torch.testing.assert_close(fused, plain, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("rows,k,n", [(4096, 64, 64), (16384, 128, 128), (5000, 134, 128), (8192, 9, 64), (2048, 3, 32),
                                      (1024, 2048, 128), (4096, 259, 256), (3000, 512, 1024), (777, 70, 48), (600, 131, 40),
                                      # >= 2 row tiles per SM with small weights: the persistent weights-resident kernel
                                      (40000, 128, 128), (50001, 134, 128), (65536, 9, 64), (38017, 3, 32),
                                      (45000, 128, 134), (40000, 64, 200), (39000, 144, 16), (41000, 70, 48), (40001, 100, 150)])
def test_tcgen05_gemm_rows_is_fp32_accurate(cuda, rows, k, n):
    """C = A.Bt^T through pasnl_gemm_rows_tf32x3 (3xTF32 on tcgen05) against fp64, held to the same 1e-4 bar as the
    rest of the path; a plain single-pass TF32 product would miss it by ~10x."""
    from pointasnl_b200 import _ops
    torch.manual_seed(rows + k + n)
    a = torch.randn(rows, k, device=cuda) * 2.0 + 0.5
    bt = torch.randn(n, k, device=cuda)
    got = _ops.gemm_rows(a, bt)
    want = a.double() @ bt.double().t()
    scale = float(want.abs().max())
    err = float((got.double() - want).abs().max())
    # 3xTF32 products are good to ~2^-21; the TMEM accumulator truncates on every add (~3e-8 * k relative, measured),
    # which is why _ops.gemm_rows cuts k into 1024-wide slabs.  Always inside the path's 1e-4 bar.
    assert err <= max(3e-6, 3e-8 * min(k, 1024)) * scale + 1e-6, (err, scale)
    assert err <= 5e-5 * scale
    ref32 = float((torch.mm(a, bt.t()).double() - want).abs().max())      # what cuBLAS fp32 itself achieves
    print('k=%d: tcgen05 max err %.2e, cuBLAS fp32 max err %.2e (|C|max %.1f)' % (k, err, ref32, scale))
    # strided operands (a column slice of a wider tensor)
    wide = torch.randn(rows, k + 8, device=cuda)
    got2 = _ops.gemm_rows(wide[:, 4:4 + k], bt)
    torch.testing.assert_close(got2.double(), wide[:, 4:4 + k].double() @ bt.double().t(), rtol=1e-4, atol=1e-4 * scale)


@pytest.mark.parametrize("rows,m,n", [(8192, 64, 64), (20000, 128, 128), (16384, 134, 128), (65536, 9, 64), (10000, 3, 32),
                                      (8192, 259, 256), (4096, 512, 1024), (5000, 70, 48), (16384, 2048, 128),
                                      # long reductions: the warp-specialised one-CTA-per-SM kernel
                                      (100000, 128, 128), (70001, 134, 128), (50000, 64, 64), (60000, 128, 32),
                                      (40000, 70, 48), (45003, 131, 64), (524288, 9, 64),
                                      # one narrow operand: fp32 streaming kernel
                                      (100000, 3, 32), (33333, 6, 30), (8192, 9, 128), (70000, 32, 3), (5000, 16, 200),
                                      (4097, 2, 2), (300000, 16, 16), (200000, 9, 16), (50000, 13, 24)])
def test_tcgen05_wgrad_gemm(cuda, rows, m, n):
    """dW = X^T Y through pasnl_gemm_wgrad_tf32x3 (tcgen05, row slabs + reductions) against fp64."""
    from pointasnl_b200 import _ops
    torch.manual_seed(rows + m + n)
    x = torch.randn(rows, m, device=cuda) + 0.3
    y = torch.randn(rows, n, device=cuda)
    got = _ops.gemm_wgrad(x, y)
    want = x.double().t() @ y.double()
    scale = float(want.abs().max())
    err = float((got.double() - want).abs().max())
    ref32 = float((torch.mm(x.t(), y).double() - want).abs().max())
    print('rows=%d: tcgen05 wgrad max err %.2e, cuBLAS fp32 %.2e (|dW|max %.1f)' % (rows, err, ref32, scale))
    assert err <= 5e-5 * scale + 1e-5, (err, scale)


@pytest.mark.parametrize("shape", [(4, 128, 64, 134), (8, 512, 32, 9), (2, 1, 512, 131), (3, 7, 1, 5), (3, 1, 300, 70), (32, 512, 512)])
def test_maxpool_over_neighbours_forward_backward_with_ties(cuda, shape):
    from pointasnl_b200 import _ops
    torch.manual_seed(sum(shape))
    x = torch.randn(*shape)
    x[..., 0, :] = x[..., -1, :]                          # duplicated neighbours: exact ties wherever they are the max
    x = torch.round(x * 4) / 4                            # and plenty more ties
    xd = x.double().requires_grad_(True)
    want = xd.amax(dim=-2)                                # torch splits the gradient evenly among ties, like TF
    xc = x.to(cuda).requires_grad_(True)
    got = _ops.MaxPoolK.apply(xc)
    assert torch.equal(got.detach().cpu().double(), want.detach())
    g = torch.randn(want.shape)
    want.backward(g.double())
    got.backward(g.to(cuda))
    torch.testing.assert_close(xc.grad.cpu().double(), xd.grad, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("shape", [(4, 128, 64, 128), (8, 512, 32, 64), (2, 40, 32, 512), (1, 3, 16, 36), (2, 10, 32, 20)])
def test_pointconv_contraction_forward_backward(cuda, shape):
    from pointasnl_b200 import _ops
    torch.manual_seed(sum(shape))
    x = torch.randn(*shape)
    w = torch.randn(*shape[:-1], 32)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    want = torch.matmul(xd.transpose(-1, -2), wd)
    xc, wc = x.to(cuda).requires_grad_(True), w.to(cuda).requires_grad_(True)
    got = _ops.pointconv_contract(xc, wc)
    torch.testing.assert_close(got.detach().cpu().double(), want.detach(), rtol=1e-4, atol=1e-5)
    g = torch.randn(want.shape)
    want.backward(g.double())
    got.backward(g.to(cuda))
    # fp32 dot products of k / 32 / c terms with O(1) inputs: the absolute floor scales with sqrt(#terms)
    torch.testing.assert_close(xc.grad.cpu().double(), xd.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(wc.grad.cpu().double(), wd.grad, rtol=1e-4, atol=2e-4)
