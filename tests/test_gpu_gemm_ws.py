"""The TMA-fed, weights-stationary rows GEMM (csrc/gemm_ws.cu: pasnl_gemm_rows_ws) against an fp64 product:
both weight layouts, the fused column statistics, the operand prologue f(a) = act(a*scale + shift), ragged row counts,
k / n that are not multiples of the tile, and the routing of pasnl_gemm_rows_tf32x3 onto it."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_ws(a, w, w_kn, stats=False, pro=None, act=0, ldc=None):
    from pointasnl_b200 import _lib, _ops
    rows, k = a.shape
    n = w.shape[1] if w_kn else w.shape[0]
    ldc = ldc or n
    c = torch.full((rows, ldc), 7.0, device=a.device)
    st = torch.zeros(2 * n, dtype=torch.float64, device=a.device) if stats else None
    sc, sh = (pro if pro is not None else (None, None))
    lib = _lib.load()
    assert (lib.pasnl_gemm_rows_ws_eligible(rows, k, n, a.data_ptr(), a.stride(0)) == 1
            or lib.pasnl_gemm_rows_ws_wide_eligible(rows, k, n, a.data_ptr(), a.stride(0)) == 1)
    _lib.call("pasnl_gemm_rows_ws", rows, k, n, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), 1 if w_kn else 0,
              c.data_ptr(), ldc, _ops._ptr(st), _ops._ptr(sc), _ops._ptr(sh), act, _ops._stream())
    return c, st


@pytest.mark.parametrize("rows,k,n", [(40000, 128, 128), (37888, 64, 64), (50001, 36, 20), (41111, 100, 70),
                                      (39000, 20, 128), (262144, 128, 128), (38000, 68, 17)])
@pytest.mark.parametrize("w_kn", [False, True])
def test_rows_ws_matches_fp64(cuda, rows, k, n, w_kn):
    torch.manual_seed(rows + k)
    a = torch.randn(rows, k, device=cuda) * 1.3 + 0.2
    w = torch.randn(k, n, device=cuda) / k ** 0.5
    wt = w if w_kn else w.t().contiguous()
    c, st = run_ws(a, wt, w_kn, stats=True)
    want = a.double() @ w.double()
    scale = float(want.abs().max())
    err = float((c.double() - want).abs().max()) / scale
    print(f"   rows={rows} k={k} n={n} w_kn={w_kn}: max error {err:.2e} of the largest output")
    assert err < 5e-6
    # the statistics are those of the fp32 result the kernel wrote (what the batch-norm pass that follows normalises)
    cd = c.double()
    # (fp32 partial sums per 128-row tile, fp64 across tiles: ~1e-7 of sum |x|)
    torch.testing.assert_close(st[:n], cd.sum(0), rtol=1e-7, atol=2e-7 * scale * rows)
    torch.testing.assert_close(st[n:], (cd * cd).sum(0), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("rows,k,n", [(16384, 128, 2048), (4100, 64, 2500), (9000, 100, 1028)])
@pytest.mark.parametrize("w_kn", [False, True])
def test_rows_ws_wide_column_blocks_match_fp64(cuda, rows, k, n, w_kn):
    """n > 128: one 128-column weight block per CTA (blockIdx.y), ragged last block and row tile, fused statistics."""
    torch.manual_seed(rows + n)
    a = torch.randn(rows, k, device=cuda) * 1.1 - 0.1
    w = torch.randn(k, n, device=cuda) / k ** 0.5
    wt = w if w_kn else w.t().contiguous()
    c, st = run_ws(a, wt, w_kn, stats=True, ldc=n + 4)
    want = a.double() @ w.double()
    scale = float(want.abs().max())
    assert float((c[:, :n].double() - want).abs().max()) / scale < 5e-6
    assert bool((c[:, n:] == 7.0).all())
    cd = c[:, :n].double()
    torch.testing.assert_close(st[:n], cd.sum(0), rtol=1e-7, atol=2e-7 * scale * rows)
    torch.testing.assert_close(st[n:], (cd * cd).sum(0), rtol=1e-6, atol=1e-9)
    # and the router sends the shape here
    from pointasnl_b200 import _ops
    got = _ops.gemm_rows(a, wt, w_kn=w_kn)
    assert torch.equal(got, c[:, :n].contiguous())


@pytest.mark.parametrize("act", [0, 1, 2])
def test_rows_ws_prologue_and_padded_output_pitch(cuda, act):
    torch.manual_seed(act)
    rows, k, n = 45000, 96, 72
    a = torch.randn(rows, k, device=cuda)
    w = torch.randn(k, n, device=cuda) / k ** 0.5
    sc, sh = torch.rand(k, device=cuda) + 0.5, torch.randn(k, device=cuda) * 0.3
    sc[::5] *= -1.0                                    # negative batch-norm scales happen
    c, _ = run_ws(a, w, True, pro=(sc, sh), act=act, ldc=80)
    z = a.double() * sc.double() + sh.double()
    z = torch.relu(z) if act == 1 else (torch.nn.functional.leaky_relu(z, 0.2) if act == 2 else z)
    # the kernel rounds z to fp32 (one fma) before splitting it: compare against the fp32-rounded operand
    want = z.float().double() @ w.double()
    err = float((c[:, :n].double() - want).abs().max()) / float(want.abs().max())
    assert err < 5e-6, err
    assert bool((c[:, n:] == 7.0).all())               # the pitch padding is not touched


def test_rows_tf32x3_routes_tall_layers_to_ws_and_both_kernels_agree(cuda, monkeypatch):
    from pointasnl_b200 import _ops
    torch.manual_seed(3)
    a = torch.randn(300000, 64, device=cuda)
    bt = torch.randn(64, 64, device=cuda) / 8
    got = _ops.gemm_rows(a, bt)
    want = a.double() @ bt.double().t()
    assert float((got.double() - want).abs().max()) / float(want.abs().max()) < 5e-6


@pytest.mark.parametrize("rows,k,n", [(16384, 1024, 128), (4096, 1024, 512), (8192, 4096, 256), (7000, 2112, 200),
                                      (12700, 516, 130), (16384, 2048, 64), (3000, 512, 1024)])
def test_rows_tma_matches_fp64_and_fuses_statistics_for_any_k(cuda, rows, k, n):
    """csrc/gemm_tma.cu through the routed entry point (pasnl_gemm_rows_tf32x3 -> pasnl_gemm_rows_tma)."""
    from pointasnl_b200 import _lib, _ops
    torch.manual_seed(rows + k + n)
    a = torch.randn(rows, k, device=cuda) * 1.5 + 0.3
    bt = torch.randn(n, k, device=cuda) / k ** 0.5
    assert _lib.load().pasnl_gemm_rows_tma_eligible(rows, k, n, a.data_ptr(), k, bt.data_ptr(), k) == 1
    st = torch.zeros(2 * n, dtype=torch.float64, device=cuda)
    c, fused = _ops.gemm_rows(a, bt, col_stats=st)
    assert fused
    want = a.double() @ bt.double().t()
    scale = float(want.abs().max())
    err = float((c.double() - want).abs().max()) / scale
    print(f"   rows={rows} k={k} n={n}: max error {err:.2e} of the largest output")
    assert err <= max(3e-6, 3e-8 * min(k, 1024))
    cd = c.double()
    torch.testing.assert_close(st[:n], cd.sum(0), rtol=1e-7, atol=2e-7 * scale * rows)
    torch.testing.assert_close(st[n:], (cd * cd).sum(0), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("rows,m,n", [(4096, 128, 128), (262144, 128, 128), (524288, 64, 64), (16384, 2048, 128),
                                      (4096, 256, 4096), (5001, 100, 72), (2048, 32, 32), (70000, 132, 96), (9999, 260, 40)])
def test_wgrad_tma_mn_major_matches_fp64(cuda, rows, m, n):
    """csrc/gemm_wgrad_tma.cu: dW = X^T Y with both operands as MN-major TMA tiles (SWIZZLE_128B, 32-byte atoms)."""
    from pointasnl_b200 import _lib, _ops
    torch.manual_seed(rows + m + n)
    x = torch.randn(rows, m, device=cuda) * 1.2 + 0.1
    dy = torch.randn(rows, n, device=cuda)
    assert _lib.load().pasnl_gemm_wgrad_tma_eligible(rows, m, n, x.data_ptr(), m, dy.data_ptr(), n) == 1
    dw = torch.empty(m, n, device=cuda)
    _lib.call("pasnl_gemm_wgrad_tma", rows, m, n, x.data_ptr(), m, dy.data_ptr(), n, dw.data_ptr(), n, 0, None, None, 0,
              _ops._stream())
    want = x.double().t() @ dy.double()
    err = float((dw.double() - want).abs().max()) / float(want.abs().max())
    print(f"   rows={rows} m={m} n={n}: max error {err:.2e} of the largest entry")
    assert err < 2e-5
    routed = _ops.gemm_wgrad(x, dy)
    assert float((routed.double() - want).abs().max()) / float(want.abs().max()) < 2e-5


def test_tall_dgrad_with_134_columns_splits_onto_the_ws_kernel(cuda):
    from pointasnl_b200 import _ops
    torch.manual_seed(9)
    dy = torch.randn(262144, 128, device=cuda)
    w = torch.randn(134, 128, device=cuda) / 11          # stored [cin=134][cout=128] = the [n][k] view of dx = dy . W^T
    dx = _ops.gemm_rows(dy, w, w_kn=False)
    want = dy.double() @ w.double().t()
    assert tuple(dx.shape) == (262144, 134)
    assert float((dx.double() - want).abs().max()) / float(want.abs().max()) < 5e-6


@pytest.mark.parametrize("rows,cin,c1,c2,act", [(40960, 9, 64, 64, 1), (38000, 134, 128, 128, 2), (40960, 64, 128, 64, 1),
                                                 (3000, 16, 64, 64, 1)])
def test_dense_bn_act_pair_matches_two_layers(rows, cin, c1, c2, act):
    """DenseBNAct2 (conv0's activation folded into conv1's GEMM / weight-gradient operand prologue) against two
    DenseBNAct calls: outputs, every gradient and the moving statistics.  (3000 rows: the unfused fallback.)"""
    from pointasnl_b200 import _ops
    g = torch.Generator(device="cpu").manual_seed(rows + c1)
    dev = "cuda"
    x = torch.randn(rows, cin, generator=g).to(dev).requires_grad_(True)

    def params():
        gg = torch.Generator(device="cpu").manual_seed(5)
        out = []
        k = cin
        for c in (c1, c2):
            out += [(torch.randn(k, c, generator=gg) / k ** 0.5).to(dev).requires_grad_(True),
                    torch.zeros(c, device=dev, requires_grad=True),
                    (1 + 0.3 * torch.randn(c, generator=gg)).to(dev).requires_grad_(True),
                    (0.3 * torch.randn(c, generator=gg)).to(dev).requires_grad_(True),
                    torch.zeros(c, device=dev), torch.ones(c, device=dev)]
            k = c
        return out
    dout = torch.randn(rows, c2, generator=g).to(dev)
    pa, pb = params(), params()
    h = _ops.DenseBNAct.apply(x, *pa[:6], 0.9, 1e-3, act, True)
    ref = _ops.DenseBNAct.apply(h, *pa[6:], 0.9, 1e-3, act, True)
    ref.backward(dout)
    gx_ref = x.grad.clone()
    x.grad = None
    out = _ops.DenseBNAct2.apply(x, *pb, 0.9, 1e-3, act, act, True)
    out.backward(dout)
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-5 * scale
    assert (x.grad - gx_ref).abs().max().item() <= 1e-4 * gx_ref.abs().max().item()
    for a, b in zip(pa, pb):
        if a.requires_grad:
            assert (a.grad - b.grad).abs().max().item() <= 1e-4 * max(a.grad.abs().max().item(), 1e-3)
        else:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("b,s,k,cin,c1,c2", [(8, 160, 32, 9, 64, 64), (4, 320, 32, 134, 128, 128)])
def test_pre_norm_pointconv_matches_materialised(b, s, k, cin, c1, c2):
    """DenseBNAct2(raw_out=True) + BNActPointConv (conv1's batch norm + ReLU applied to the tile inside the PointConv
    kernels, backward continued through the batch norm) against DenseBNAct2 + PointConvContract."""
    from pointasnl_b200 import _ops
    g = torch.Generator(device="cpu").manual_seed(b + s)
    dev = "cuda"
    x = torch.randn(b, s, k, cin, generator=g).to(dev).requires_grad_(True)
    wn = torch.randn(b, s, k, 32, generator=g).to(dev).requires_grad_(True)
    assert _ops.pair_plan(x.reshape(-1, cin), c1, c2) and _ops.pointconv_eligible(k, c2)

    def params():
        gg = torch.Generator(device="cpu").manual_seed(5)
        out, kk = [], cin
        for c in (c1, c2):
            out += [(torch.randn(kk, c, generator=gg) / kk ** 0.5).to(dev).requires_grad_(True),
                    torch.zeros(c, device=dev, requires_grad=True),
                    (1 + 0.3 * torch.randn(c, generator=gg)).to(dev).requires_grad_(True),
                    (0.3 * torch.randn(c, generator=gg)).to(dev).requires_grad_(True),
                    torch.zeros(c, device=dev), torch.ones(c, device=dev)]
            kk = c
        return out
    dout = torch.randn(b, s, c2, 32, generator=g).to(dev)
    pa, pb = params(), params()
    h = _ops.DenseBNAct2.apply(x, *pa, 0.9, 1e-3, 1, 1, True)
    ref = _ops.PointConvContract.apply(h, wn)
    ref.backward(dout)
    gx_ref, gw_ref = x.grad.clone(), wn.grad.clone()
    x.grad = wn.grad = None
    y, mean2, var2, scale2, shift2 = _ops.DenseBNAct2.apply(x, *pb, 0.9, 1e-3, 1, 1, True, True)
    out = _ops.BNActPointConv.apply(y, pb[8], pb[9], mean2, var2, scale2, shift2, wn, 1e-3, 1)
    out.backward(dout)
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert (x.grad - gx_ref).abs().max().item() <= 1e-4 * gx_ref.abs().max().item()
    assert (wn.grad - gw_ref).abs().max().item() <= 1e-4 * gw_ref.abs().max().item()
    for a, c in zip(pa, pb):
        if a.requires_grad:
            assert (a.grad - c.grad).abs().max().item() <= 1e-4 * max(a.grad.abs().max().item(), 1e-3)
        else:
            assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)
