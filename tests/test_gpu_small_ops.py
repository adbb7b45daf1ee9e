"""csrc/small_ops.cu against plain PyTorch: skinny GEMMs of the classifier head, add / concat / split, dropout,
soft-max cross entropy (both reductions), the fused SA-layer front end (max-pool scatter gradient with ties, slice
gradient added in place), bias add / column sum, the batched weight transpose and the flat L2 + Adam step (against a
numpy restatement of tf.train.AdamOptimizer's update)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,k,n", [(32, 1536, 512), (32, 512, 256), (32, 256, 40), (64, 300, 70), (5, 7, 3), (1, 129, 65)])
def test_skinny_gemms_forward_dgrad_wgrad(cuda, rows, k, n):
    from pointasnl_b200 import _ops
    torch.manual_seed(rows + k)
    x = torch.randn(rows, k, device=cuda)
    w = torch.randn(k, n, device=cuda) / k ** 0.5
    dy = torch.randn(rows, n, device=cuda)
    y = _ops.gemm_rows(x, w, w_kn=True)
    torch.testing.assert_close(y.double(), x.double() @ w.double(), rtol=1e-5, atol=1e-5)
    y2 = _ops.gemm_rows(x, w.t().contiguous(), w_kn=False)
    torch.testing.assert_close(y2.double(), x.double() @ w.double(), rtol=1e-5, atol=1e-5)
    dx = _ops.gemm_rows(dy, w, w_kn=False)                  # dy . W^T with W stored [k][n] seen as [n'=k][k'=n]
    torch.testing.assert_close(dx.double(), dy.double() @ w.double().t(), rtol=1e-5, atol=1e-5)
    dw = _ops.gemm_wgrad(x, dy)
    torch.testing.assert_close(dw.double(), x.double().t() @ dy.double(), rtol=1e-5, atol=1e-5)


def test_dense_bias_layers_and_their_gradients(cuda):
    from pointasnl_b200 import _ops
    torch.manual_seed(0)
    for rows, cin, cout in ((32, 256, 40), (40960, 128, 21), (300, 64, 13)):
        x = torch.randn(rows, cin, device=cuda, requires_grad=True)
        w = (torch.randn(cin, cout, device=cuda) / cin ** 0.5).requires_grad_(True)
        b = torch.randn(cout, device=cuda, requires_grad=True)
        out = _ops.DenseBias.apply(x, w, b)
        g = torch.randn_like(out)
        out.backward(g)
        xd, wd, bd = [t.detach().double().requires_grad_(True) for t in (x, w, b)]
        want = xd @ wd + bd
        want.backward(g.double())
        torch.testing.assert_close(out.detach().double(), want.detach(), rtol=1e-5, atol=2e-5)
        for got, ref in ((x.grad, xd.grad), (w.grad, wd.grad), (b.grad, bd.grad)):
            assert float((got.double() - ref).norm() / ref.norm()) < 1e-5


def test_add_concat_split(cuda):
    from pointasnl_b200 import _ops
    torch.manual_seed(1)
    a, b, c = [torch.randn(7, 33, 5, device=cuda, requires_grad=True) for _ in range(3)]
    out = _ops.add_n(a, b, c)
    assert torch.equal(out, (a + b) + c)
    out.backward(torch.ones_like(out))
    assert torch.equal(a.grad, torch.ones_like(a)) and torch.equal(c.grad, torch.ones_like(c))
    assert torch.equal(_ops.add_n(a.detach(), b.detach()), a.detach() + b.detach())
    p, q = torch.randn(4, 9, 3, device=cuda, requires_grad=True), torch.randn(4, 9, 131, device=cuda, requires_grad=True)
    cat = _ops.concat_cols(p, q)
    assert torch.equal(cat, torch.cat([p, q], -1))
    g = torch.randn_like(cat)
    cat.backward(g)
    assert torch.equal(p.grad, g[..., :3]) and torch.equal(q.grad, g[..., 3:])
    x = torch.randn(3, 50, 6, device=cuda)
    u, v = _ops.split_cols(x, 3)
    assert torch.equal(u, x[..., :3]) and torch.equal(v, x[..., 3:]) and u.is_contiguous()


def test_dropout_mask_rate_scaling_and_backward_consistency(cuda):
    from pointasnl_b200 import _ops
    step = torch.zeros((), dtype=torch.int64, device=cuda)
    _ops.set_dropout_step(step)
    try:
        x = torch.ones(200000, device=cuda, requires_grad=True)
        y = _ops.dropout(x, 0.4, 1234)
        kept = (y != 0)
        assert abs(float(kept.float().mean()) - 0.4) < 0.01
        assert torch.allclose(y[kept], torch.full_like(y[kept], 1 / 0.4))
        y.backward(torch.full_like(y, 3.0))
        assert torch.equal(x.grad != 0, kept) and torch.allclose(x.grad[kept], torch.full_like(y[kept], 3.0 / 0.4))
        y2 = _ops.dropout(x.detach(), 0.4, 1234)
        assert torch.equal(y2, y.detach())                      # same (seed, step) -> same mask
        step += 1
        y3 = _ops.dropout(x.detach(), 0.4, 1234)
        assert not torch.equal(y3, y.detach())                  # the optimizer step advances the stream
        assert not torch.equal(_ops.dropout(x.detach(), 0.4, 99), y3)
    finally:
        _ops.set_dropout_step(None)


@pytest.mark.parametrize("weighted", [False, True])
def test_softmax_cross_entropy_value_and_gradient(cuda, weighted):
    from pointasnl_b200 import _ops
    torch.manual_seed(2)
    rows, c = (4 * 1000, 21) if weighted else (32, 40)
    logits = (torch.randn(rows, c, device=cuda) * 3).requires_grad_(True)
    labels = torch.randint(0, c, (rows,), device=cuda)
    w = None
    if weighted:
        w = torch.rand(rows, device=cuda) * 4 + 1
        w[::7] = 0.0
    loss = _ops.softmax_ce(logits.reshape(4, -1, c) if weighted else logits, labels, w, out_scale=2.0 if weighted else 1.0)
    loss.backward(torch.tensor(0.5, device=cuda))
    ld = logits.detach().double().requires_grad_(True)
    ce = torch.nn.functional.cross_entropy(ld, labels, reduction="none")
    want = 2.0 * (ce * w.double()).sum() / torch.count_nonzero(w) if weighted else ce.mean()
    (0.5 * want).backward()
    assert abs(float(loss) - float(want)) <= 1e-6 * abs(float(want))
    assert float((logits.grad.double() - ld.grad).abs().max()) <= 1e-6 * float(ld.grad.abs().max()) + 1e-9


def test_sa_front_matches_autograd_of_the_three_readers_with_ties(cuda):
    from pointasnl_b200 import _ops
    torch.manual_seed(3)
    b, s, k, w = 2, 37, 16, 11
    base = torch.randn(b, s, k, w, device=cuda)
    base[:, :, 3] = base[:, :, 5]                       # duplicated neighbours -> exact ties in every channel
    base[0, 0, :, 2] = 1.25                             # a channel that is constant over the whole neighbourhood
    x = base.clone().requires_grad_(True)
    pooled, ident, xyz = _ops.SAFront.apply(x)
    wt = torch.randn(w, 6, device=cuda)
    g1, g2, g3 = torch.randn(b, s, w, device=cuda), torch.randn(b, s, k, 6, device=cuda), torch.randn(b, s, k, 3, device=cuda)
    ((pooled * g1).sum() + ((ident @ wt) * g2).sum() + (xyz * g3).sum()).backward()
    xr = base.clone().double().requires_grad_(True)
    pr = xr.amax(dim=2)                                  # torch.amax splits the gradient evenly among ties, like TF
    ((pr * g1.double()).sum() + ((xr @ wt.double()) * g2.double()).sum() + (xr[..., :3] * g3.double()).sum()).backward()
    assert torch.equal(pooled, base.amax(dim=2)) and torch.equal(xyz, base[..., :3]) and torch.equal(ident, base)
    torch.testing.assert_close(x.grad.double(), xr.grad, rtol=1e-5, atol=1e-5)


def test_bias_add_col_sum_and_batched_transpose(cuda):
    from pointasnl_b200 import _lib, _ops
    torch.manual_seed(4)
    ws = [torch.randn(*shape, device=cuda) for shape in ((1, 1, 134, 128), (259, 256), (1, 64, 32, 128), (3, 5))]
    tr = _ops.WeightTransposer()
    tr.register(ws)
    tr.refresh()
    for w in ws:
        n = w.shape[-1]
        w2 = w.reshape(-1, n)
        assert torch.equal(tr.lookup(w2), w2.t().contiguous())
    w2 = torch.randn(70, 33, device=cuda)
    assert torch.equal(_ops.transposed_weight(w2), w2.t().contiguous())     # unregistered: single-matrix launch
    x = torch.randn(5000, 37, device=cuda)
    s = torch.empty(37, device=cuda)
    _lib.call("pasnl_col_sum", 5000, 37, x.data_ptr(), s.data_ptr(), _ops._stream())
    torch.testing.assert_close(s.double(), x.double().sum(0), rtol=1e-5, atol=1e-4)


def test_flat_adam_with_l2_follows_tf_adam(cuda):
    """tf.train.AdamOptimizer: m, v moments; lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps); the L2 term
    1e-4 * p enters the gradient of the masked ('weights') elements; the step counter and the loss are published."""
    from pointasnl_b200 import _ops
    rng = np.random.default_rng(5)
    n = 10007
    p0 = rng.normal(size=n).astype(np.float32)
    mask = (rng.uniform(size=n) < 0.6)
    p = torch.from_numpy(p0.copy()).to(cuda)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros((), dtype=torch.int64, device=cuda)
    reg = torch.zeros((), dtype=torch.float64, device=cuda)
    loss_in, loss_out = torch.tensor(1.5, device=cuda), torch.zeros((), device=cuda)
    pm = p0.astype(np.float64)
    mm, vm = np.zeros(n), np.zeros(n)
    lr, b1, b2, eps, l2 = 1e-3, 0.9, 0.999, 1e-8, 1e-4
    for t in range(1, 4):
        g = rng.normal(size=n).astype(np.float32)
        _ops.adam_l2(p, torch.from_numpy(g).to(cuda), m, v, torch.from_numpy(mask.astype(np.uint8)).to(cuda), l2, 0.5, lr,
                     step, reg, loss_in, loss_out)
        want_reg = 0.5 * l2 * (pm[mask] ** 2).sum()
        gt = 0.5 * g.astype(np.float64) + l2 * pm * mask
        mm = b1 * mm + (1 - b1) * gt
        vm = b2 * vm + (1 - b2) * gt * gt
        pm = pm - lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t) * mm / (np.sqrt(vm) + eps)
        assert int(step) == t
        assert abs(float(loss_out) - (1.5 + want_reg)) <= 1e-5 * (1.5 + want_reg)
        assert float(reg) == 0.0
    np.testing.assert_allclose(p.cpu().numpy(), pm, rtol=2e-5, atol=2e-6)
