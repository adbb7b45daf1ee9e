"""GPU parity tests proper (run on the B200 box: pytest -m gpu).

Every test drives the CUDA path THROUGH THE C ABI (pointasnl_b200._ops -> libpasnl_b200.so) and compares it with
  (1) the CPU oracle (oracle/cpu_ref.c, oracle/torch_ref.py), and where it exists
  (2) the reference's own code compiled untouched (oracle/_ref: GPU kernels for FPS / gather / ball query / group,
      nanoflann kNN and the three_nn / three_interpolate CPU functions).
Bars: bit-exact (torch.equal) for integer / index / copy work and for three_interpolate's forward; rtol 1e-4 for the
floating-point attention cores and the scatter-add gradients (atomic order), as BASELINE.json's north_star states.
"""
import numpy as np
import pytest
import torch

from oracle import cref, refnative, torch_ref
from tests import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: "fp32 attention/interpolation within 1e-4 rel"


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def have_ref(name):
    return refnative.available(name)


# ------------------------------------------------------------------ FPS
FPS_CASES = [
    ("cfg1", lambda: synth.uniform_clouds(2, 1024, 0), 512),
    ("cls_l1", lambda: synth.modelnet_batch(32, 1024, 1)[0][:, :, :3], 512),
    ("cls_l2", lambda: synth.uniform_clouds(32, 512, 5), 128),
    ("lattice", lambda: synth.lattice_clouds(3, 700, 4, 1), 300),
    ("lattice_big", lambda: synth.lattice_clouds(2, 3073, 6, 2), 400),
    ("dups", lambda: synth.duplicated_clouds(4, 1024, 0.05, 3), 512),
    ("n40", lambda: synth.uniform_clouds(4, 40, 6), 32),
    ("n80", lambda: synth.uniform_clouds(4, 80, 7), 40),
    ("n513", lambda: synth.uniform_clouds(3, 513, 8), 200),
    ("n3072", lambda: synth.uniform_clouds(2, 3072, 9), 300),
    ("n5000", lambda: synth.uniform_clouds(2, 5000, 10), 300),
    ("m_gt_n", lambda: synth.uniform_clouds(2, 100, 11), 130),
    ("scannet", lambda: synth.scannet_batch(2, 10240, 2, 0.05)[0][:, :, :3], 1280),  # cluster of 2 CTAs
    ("kitti", lambda: synth.kitti_batch(1, 65536, 3)[0][:, :, :3], 512),            # cluster of 8 CTAs
    ("huge", lambda: synth.uniform_clouds(1, 70001, 12, normalize=False), 64),      # global-scratch fallback
]


@pytest.mark.parametrize("name,gen,m", FPS_CASES, ids=[c[0] for c in FPS_CASES])
def test_fps_bit_exact(cuda, name, gen, m):
    from pointasnl_b200 import tf_sampling
    xyz = gen()
    got = tf_sampling.farthest_point_sample(m, dev(xyz, cuda))
    assert got.dtype == torch.int32 and tuple(got.shape) == (xyz.shape[0], m)
    want = cref.fps(xyz, m)
    assert np.array_equal(got.cpu().numpy(), want), f"FPS differs from the CPU oracle on {name}"
    if have_ref("ref_sampling.so"):
        ref = refnative.gpu_fps(dev(xyz, cuda), m)
        assert torch.equal(got, ref), f"FPS differs from the compiled reference kernel on {name}"


def test_gather_point_and_grad(cuda):
    from pointasnl_b200 import tf_sampling
    xyz = synth.uniform_clouds(8, 1024, 0)
    x = dev(xyz, cuda).requires_grad_(True)
    idx = tf_sampling.farthest_point_sample(256, x.detach())
    out = tf_sampling.gather_point(x, idx)
    assert np.array_equal(out.detach().cpu().numpy(), cref.gather_point(xyz, idx.cpu().numpy()))
    g = torch.randn_like(out)
    out.backward(g)
    want = cref.gather_point_grad(1024, idx.cpu().numpy(), g.cpu().numpy())
    torch.testing.assert_close(x.grad.cpu(), torch.from_numpy(want), rtol=RTOL, atol=1e-6)
    if have_ref("ref_sampling.so"):
        assert torch.equal(out.detach(), refnative.gpu_gather_point(x.detach(), idx))
        torch.testing.assert_close(x.grad, refnative.gpu_gather_point_grad(1024, idx, g), rtol=RTOL, atol=1e-6)
    # duplicate indices accumulate
    idx2 = torch.zeros((8, 16), dtype=torch.int32, device=cuda)
    x2 = dev(xyz, cuda).requires_grad_(True)
    tf_sampling.gather_point(x2, idx2).sum().backward()
    assert torch.allclose(x2.grad[:, 0], torch.full((8, 3), 16.0, device=cuda)) and x2.grad[:, 1:].abs().sum() == 0


# ------------------------------------------------------------------ ball query
def _ball_case(cuda, xyz1, xyz2, radius, nsample):
    from pointasnl_b200 import tf_grouping
    idx, cnt = tf_grouping.query_ball_point(radius, nsample, dev(xyz1, cuda), dev(xyz2, cuda))
    wi, wc = cref.ball_query(radius, nsample, xyz1, xyz2)
    assert np.array_equal(cnt.cpu().numpy(), wc)
    assert np.array_equal(idx.cpu().numpy(), wi)
    if have_ref("ref_grouping.so"):
        ri, rc = refnative.gpu_ball_query(radius, nsample, dev(xyz1, cuda), dev(xyz2, cuda))
        assert torch.equal(cnt, rc)
        assert torch.equal(idx, ri)
    return idx, cnt


def test_ball_query_config1(cuda):
    xyz = synth.uniform_clouds(2, 1024, 0)
    q = cref.gather_point(xyz, cref.fps(xyz, 512))
    _ball_case(cuda, xyz, q, 0.2, 32)      # queries = FPS output
    _ball_case(cuda, xyz, xyz, 0.2, 32)    # self query


def test_ball_query_reference_smoke_shape(cuda):
    # the only shape the reference authors exercised in isolation (tf_grouping.py:80-88)
    rng = np.random.default_rng(100)
    x1 = rng.random((32, 512, 3)).astype(np.float32)
    x2 = rng.random((32, 128, 3)).astype(np.float32)
    _ball_case(cuda, x1, x2, 0.1, 64)


def test_ball_query_edges(cuda):
    lat = synth.lattice_clouds(2, 900, 5, 3)
    _ball_case(cuda, lat, lat[:, :300], 1.0, 16)          # d == radius exactly is a miss (strict '<')
    _ball_case(cuda, lat, lat[:, :300], float(np.sqrt(np.float32(2.0))), 16)
    _ball_case(cuda, lat, lat[:, :300], 1.0000001, 8)
    xyz = synth.uniform_clouds(2, 5000, 4)                # several shared-memory tiles
    _ball_case(cuda, xyz, xyz[:, ::7], 0.15, 20)
    idx, cnt = _ball_case(cuda, xyz, xyz[:, :64] + 100.0, 0.1, 8)   # no hit at all: cnt 0, idx defined as 0
    assert int(cnt.sum()) == 0 and int(idx.abs().sum()) == 0
    # repulsion-loss shape (pointasnl_util.py:361-363): S = N = 1280, K = 20, r = 0.07
    sc = synth.scannet_batch(2, 1280, 2)[0][:, :, :3]
    _ball_case(cuda, sc, sc, 0.07, 20)


# ------------------------------------------------------------------ group_point
@pytest.mark.parametrize("c", [3, 6, 64, 131])
def test_group_point_and_grad(cuda, c):
    from pointasnl_b200 import tf_grouping
    rng = np.random.default_rng(c)
    b, n, m, ns = 4, 512, 128, 32
    pts = rng.normal(size=(b, n, c)).astype(np.float32)
    idx = rng.integers(0, n, (b, m, ns)).astype(np.int32)
    p = dev(pts, cuda).requires_grad_(True)
    i = dev(idx, cuda)
    out = tf_grouping.group_point(p, i)
    assert np.array_equal(out.detach().cpu().numpy(), cref.group_point(pts, idx))
    g = torch.randn_like(out)
    out.backward(g)
    want = cref.group_point_grad(n, idx, g.cpu().numpy())
    torch.testing.assert_close(p.grad.cpu(), torch.from_numpy(want), rtol=RTOL, atol=1e-5)
    if have_ref("ref_grouping.so"):
        assert torch.equal(out.detach(), refnative.gpu_group_point(p.detach(), i))
        torch.testing.assert_close(p.grad, refnative.gpu_group_point_grad(n, i, g), rtol=RTOL, atol=1e-5)


# ------------------------------------------------------------------ k-NN
KNN_CASES = [
    ("cls_l1", 4, 1024, 512, 32),
    ("cls_l2", 4, 512, 128, 64),
    ("k20", 2, 700, 300, 20),
    ("k128", 2, 600, 100, 128),
    ("n40", 4, 40, 40, 32),       # seg layer4_2: K is 80 % of the cloud
    ("tiles", 1, 5000, 900, 32),  # several shared-memory tiles
]


@pytest.mark.parametrize("name,b,n,m,k", KNN_CASES, ids=[c[0] for c in KNN_CASES])
def test_knn_matches_oracle_and_nanoflann(cuda, name, b, n, m, k):
    from pointasnl_b200 import pointasnl_util, tf_grouping
    xyz = synth.uniform_clouds(b, n, 20 + n)
    q = xyz[:, :m] if m <= n else synth.uniform_clouds(b, m, 3)
    val, idx = tf_grouping.knn_point(k, dev(xyz, cuda), dev(q, cuda))
    wd, wi = cref.knn(k, xyz, q, return_dist=True)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(val.cpu().numpy(), wd)          # distances bit-exact: same no-FMA arithmetic
    assert np.array_equal(idx[:, :, 0].cpu().numpy(), np.broadcast_to(np.arange(m, dtype=np.int32), (b, m)))  # self first
    if have_ref("ref_knn.so"):
        ref = refnative.knn_batch(xyz, q, k)               # the kd-tree the reference models call
        same = ref == wi
        if not same.all():
            # only exact distance ties may differ in order (nanoflann keeps tree-visit order)
            bb, mm, kk = np.nonzero(~same)
            d_ref = ((xyz[bb, ref[bb, mm, kk]] - q[bb, mm]) ** 2)
            d_ref = (d_ref[:, 0] + d_ref[:, 1]) + d_ref[:, 2]
            assert np.array_equal(d_ref.astype(np.float32), wd[bb, mm, kk])
    got2 = pointasnl_util.knn_query(k, dev(xyz, cuda), dev(q, cuda))
    assert torch.equal(got2, idx)


def test_knn_ties_follow_index_order(cuda):
    from pointasnl_b200 import tf_grouping
    lat = synth.lattice_clouds(2, 800, 4, 9)
    val, idx = tf_grouping.knn_point(32, dev(lat, cuda), dev(lat[:, :200], cuda))
    wd, wi = cref.knn(32, lat, lat[:, :200], return_dist=True)
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(val.cpu().numpy(), wd)
    # selection_sort (tf_grouping_g.cu:83-123) swaps the found minimum into place, so it is NOT stable: among
    # equidistant candidates its index order depends on the swap history.  Distances must agree position by
    # position; indices must agree wherever the distance is unique in the row and always point at that distance.
    dist = ((lat[:, :200, None, :] - lat[:, None, :, :]) ** 2)
    dist = (dist[..., 0] + dist[..., 1]) + dist[..., 2]
    oi, od = cref.selection_sort(32, dist)
    assert np.array_equal(od[:, :, :32], wd)
    assert np.array_equal(np.take_along_axis(dist, oi[:, :, :32].astype(np.int64), 2), wd)
    if have_ref("ref_grouping.so"):
        ri, rd = refnative.gpu_selection_sort(32, dev(dist, cuda))
        assert torch.equal(rd[:, :, :32], val)
        assert np.array_equal(ri.cpu().numpy(), oi) and np.array_equal(rd.cpu().numpy(), od)   # oracle == reference kernel


# ------------------------------------------------------------------ three_nn / three_interpolate
@pytest.mark.parametrize("b,n,m", [(4, 1280, 320), (2, 5000, 2500), (4, 80, 40), (3, 33, 2), (2, 512, 128)])
def test_three_nn_bit_exact(cuda, b, n, m):
    from pointasnl_b200 import tf_interpolate
    x1 = synth.uniform_clouds(b, n, 30 + n)
    x2 = x1[:, ::max(1, n // m)][:, :m].copy() if m <= n else synth.uniform_clouds(b, m, 31)
    dist, idx = tf_interpolate.three_nn(dev(x1, cuda), dev(x2, cuda))
    wd, wi = cref.three_nn(x1, x2)
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(dist.cpu().numpy(), wd)
    if have_ref("ref_interp.so"):
        rd, ri = refnative.three_nn(x1, x2)
        assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd)


def test_three_nn_ties(cuda):
    from pointasnl_b200 import tf_interpolate
    lat = synth.lattice_clouds(2, 1000, 4, 5)
    dist, idx = tf_interpolate.three_nn(dev(lat, cuda), dev(lat[:, :300], cuda))
    wd, wi = cref.three_nn(lat, lat[:, :300])
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(dist.cpu().numpy(), wd)


@pytest.mark.parametrize("c", [64, 128, 35, 1])
def test_three_interpolate_and_grad(cuda, c):
    from pointasnl_b200 import tf_interpolate
    rng = np.random.default_rng(100 + c)
    b, n, m = 4, 1280, 320
    x1 = synth.uniform_clouds(b, n, 40)
    x2 = x1[:, ::4].copy()
    pts = rng.normal(size=(b, m, c)).astype(np.float32)
    d, i = cref.three_nn(x1, x2)
    d = np.maximum(d, 1e-10)
    w = ((1.0 / d) / (1.0 / d).sum(-1, keepdims=True)).astype(np.float32)   # pointnet_util.py:212-215
    p = dev(pts, cuda).requires_grad_(True)
    out = tf_interpolate.three_interpolate(p, dev(i, cuda), dev(w, cuda))
    assert np.array_equal(out.detach().cpu().numpy(), cref.three_interpolate(pts, i, w))   # bit-exact: no FMA
    g = torch.randn_like(out)
    out.backward(g)
    want = cref.three_interpolate_grad(m, i, w, g.cpu().numpy())
    torch.testing.assert_close(p.grad.cpu(), torch.from_numpy(want), rtol=RTOL, atol=1e-5)
    if have_ref("ref_interp.so"):
        assert np.array_equal(out.detach().cpu().numpy(), refnative.three_interpolate(pts, i, w))
        torch.testing.assert_close(p.grad.cpu(), torch.from_numpy(refnative.three_interpolate_grad(m, i, w, g.cpu().numpy())),
                                   rtol=RTOL, atol=1e-5)


# ------------------------------------------------------------------ fused SA front end
@pytest.mark.parametrize("c,with_center,prefix", [(3, True, None), (128, True, None), (35, False, 8), (0, True, None), (6, True, 12)])
def test_group_concat_and_grad(cuda, c, with_center, prefix):
    from pointasnl_b200 import _ops
    rng = np.random.default_rng(7 + c)
    b, n, m, ns = 3, 400, 100, 16
    xyz = rng.normal(size=(b, n, 3)).astype(np.float32)
    feat = rng.normal(size=(b, n, c)).astype(np.float32) if c else None
    idx_full = rng.integers(0, n, (b, m, ns)).astype(np.int32)
    ctr = rng.normal(size=(b, m, 3)).astype(np.float32) if with_center else None
    x = dev(xyz, cuda).requires_grad_(True)
    f = dev(feat, cuda).requires_grad_(True) if c else None
    ce = dev(ctr, cuda).requires_grad_(True) if with_center else None
    i_full = dev(idx_full, cuda)
    i = i_full[:, :, :prefix] if prefix else i_full
    out = _ops.GroupConcat.apply(x, f, i, ce)
    idx = idx_full[:, :, :prefix] if prefix else idx_full
    gx = cref.group_point(xyz, idx)
    parts = ([gx - ctr[:, :, None, :]] if with_center else []) + [gx] + ([cref.group_point(feat, idx)] if c else [])
    want = np.concatenate(parts, -1)
    assert np.array_equal(out.detach().cpu().numpy(), want)       # copies + one IEEE subtraction: bit-exact
    # gradient against plain torch autograd of the same composition (fp64 on the CPU)
    g = torch.randn_like(out)
    out.backward(g)
    xr = torch.from_numpy(xyz).double().requires_grad_(True)
    fr = torch.from_numpy(feat).double().requires_grad_(True) if c else None
    cr = torch.from_numpy(ctr).double().requires_grad_(True) if with_center else None
    li = torch.from_numpy(idx).long()
    bi = torch.arange(b)[:, None, None]
    gxr = xr[bi, li]
    ref_parts = ([gxr - cr[:, :, None, :]] if with_center else []) + [gxr] + ([fr[bi, li]] if c else [])
    torch.cat(ref_parts, -1).backward(g.cpu().double())
    torch.testing.assert_close(x.grad.cpu().double(), xr.grad, rtol=RTOL, atol=1e-5)
    if c:
        torch.testing.assert_close(f.grad.cpu().double(), fr.grad, rtol=RTOL, atol=1e-5)
    if with_center:
        torch.testing.assert_close(ce.grad.cpu().double(), cr.grad, rtol=RTOL, atol=1e-5)


# ------------------------------------------------------------------ Adaptive Sampling cores
@pytest.mark.parametrize("groups,k,d", [((32, 64), 12, 32), ((8, 32), 12, 65), ((4, 100), 8, 32), ((4, 50), 4, 33), ((2, 5), 32, 40), ((1, 3), 1, 7)])
def test_as_core_fwd_bwd(cuda, groups, k, d):
    from pointasnl_b200 import _ops
    torch.manual_seed(k * 100 + d)
    q = torch.randn(*groups, k, d)
    kv = torch.randn(*groups, k, 2 * d)
    scale = 1.0 / np.sqrt(d)
    qd, kvd = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    want, _ = torch_ref.as_core(qd, kvd, scale)
    qc, kvc = q.to(cuda).requires_grad_(True), kv.to(cuda).requires_grad_(True)
    got = _ops.ASCore.apply(qc, kvc, scale)
    torch.testing.assert_close(got.detach().cpu().double(), want.detach(), rtol=RTOL, atol=1e-5)
    g = torch.randn_like(q)
    want.backward(g.double())
    got.backward(g.to(cuda))
    torch.testing.assert_close(qc.grad.cpu().double(), qd.grad, rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(kvc.grad.cpu().double(), kvd.grad, rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("b,s,K,k,cg", [(4, 64, 32, 12, 6), (2, 32, 64, 12, 131), (2, 40, 32, 8, 35), (2, 10, 16, 4, 67)])
def test_as_resample_fwd_bwd_strided_views(cuda, b, s, K, k, cg):
    from pointasnl_b200 import _ops
    torch.manual_seed(cg)
    grouped = torch.randn(b, s, K, 3 + cg)            # [rel xyz | ...]: AS reads first-k views of wider tensors
    gxyz = torch.randn(b, s, K, 3)
    logits = torch.randn(b, s, k, 1 + cg)
    ld = logits.double().requires_grad_(True)
    xd = gxyz.double().requires_grad_(True)
    fd = grouped.double().requires_grad_(True)
    wx, wf, _ = torch_ref.as_resample(ld, xd[:, :, :k], fd[:, :, :k, 3:])
    lc = logits.to(cuda).requires_grad_(True)
    xc = gxyz.to(cuda).requires_grad_(True)
    fc = grouped.to(cuda).requires_grad_(True)
    gx_, gf_ = _ops.ASResample.apply(lc, xc[:, :, :k], fc[:, :, :k, 3:])
    torch.testing.assert_close(gx_.detach().cpu().double(), wx.detach(), rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(gf_.detach().cpu().double(), wf.detach(), rtol=RTOL, atol=1e-5)
    g1, g2 = torch.randn(b, s, 3), torch.randn(b, s, cg)
    (wx * g1.double()).sum().add((wf * g2.double()).sum()).backward()
    (gx_ * g1.to(cuda)).sum().add((gf_ * g2.to(cuda)).sum()).backward()
    torch.testing.assert_close(lc.grad.cpu().double(), ld.grad, rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(xc.grad.cpu().double(), xd.grad, rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(fc.grad.cpu().double(), fd.grad, rtol=RTOL, atol=1e-5)


# ------------------------------------------------------------------ Non-Local core
NL_CASES = [(4, 512, 1024, 32), (4, 128, 512, 64), (2, 1280, 2000, 32), (2, 40, 80, 128), (2, 70, 130, 36), (1, 1, 5, 4), (2, 100, 33, 64),
            # d = 32: tcgen05 backward (ragged query / key tails; few query tiles -> key range split with atomics)
            (3, 200, 333, 32), (1, 130, 4100, 32), (2, 65, 129, 32)]


@pytest.mark.parametrize("b,s,n,d", NL_CASES)
def test_nl_core_fwd_bwd(cuda, b, s, n, d):
    from pointasnl_b200 import _ops
    torch.manual_seed(s + n + d)
    q = torch.randn(b, s, d) * 1.5
    kv = torch.randn(b, n, 2 * d) * 1.5
    scale = 1.0 / np.sqrt(d)
    qd, kvd = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    want = torch_ref.nl_core(qd, kvd, scale)
    qc, kvc = q.to(cuda).requires_grad_(True), kv.to(cuda).requires_grad_(True)
    got = _ops.NLCore.apply(qc, kvc, scale)
    torch.testing.assert_close(got.detach().cpu().double(), want.detach(), rtol=RTOL, atol=1e-5)
    g = torch.randn_like(q)
    want.backward(g.double())
    got.backward(g.to(cuda))
    torch.testing.assert_close(qc.grad.cpu().double(), qd.grad, rtol=RTOL, atol=2e-5)
    torch.testing.assert_close(kvc.grad.cpu().double(), kvd.grad, rtol=RTOL, atol=2e-5)


def test_ops_reject_cpu_tensors_and_bad_shapes(cuda):
    from pointasnl_b200 import tf_grouping, tf_interpolate, tf_sampling
    with pytest.raises(ValueError, match="no CPU path"):
        tf_sampling.farthest_point_sample(4, torch.zeros(1, 8, 3))
    with pytest.raises(ValueError, match="FarthestPointSample expects"):
        tf_sampling.farthest_point_sample(4, torch.zeros(1, 8, 4, device=cuda))
    with pytest.raises(ValueError, match="positive npoint"):
        tf_sampling.farthest_point_sample(0, torch.zeros(1, 8, 3, device=cuda))
    with pytest.raises(ValueError, match="QueryBallPoint expects positive radius"):
        tf_grouping.query_ball_point(0.0, 4, torch.zeros(1, 8, 3, device=cuda), torch.zeros(1, 2, 3, device=cuda))
    with pytest.raises(ValueError, match="ThreeNN expects"):
        tf_interpolate.three_nn(torch.zeros(1, 8, 2, device=cuda), torch.zeros(1, 2, 3, device=cuda))
