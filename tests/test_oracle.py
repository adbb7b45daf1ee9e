"""CPU tests of the oracle (no GPU): the C restatement against the golden vectors produced by the REFERENCE's own
code (tests/golden/*.npz, generators committed beside them), against the compiled reference when oracle/_ref is
present, and against size-independent properties."""
import os

import numpy as np
import pytest

from oracle import cref, refnative
from tests import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold_fps():
    return np.load(os.path.join(GOLD, "gpu_ref_fps.npz"))


@pytest.fixture(scope="module")
def gold_grp():
    return np.load(os.path.join(GOLD, "gpu_ref_grouping.npz"))


@pytest.fixture(scope="module")
def gold_cpu():
    return np.load(os.path.join(GOLD, "cpu_ref.npz"))


# ---- FPS: oracle == the reference GPU kernel's recorded output (bit-exact), on random, lattice, duplicate inputs
@pytest.mark.parametrize("name", ["cfg1", "lattice", "dups", "n513", "n3073"])
def test_fps_oracle_matches_reference_gpu_golden(gold_fps, name):
    xyz, want = gold_fps[f"{name}_xyz"], gold_fps[f"{name}_idx"]
    assert np.array_equal(cref.fps(xyz, want.shape[1]), want)


def test_fps_generators_are_reproducible(gold_fps):
    # the GPU parity tests regenerate inputs from seeds; they must be the very arrays the golden run saw
    assert np.array_equal(synth.uniform_clouds(2, 1024, 0), gold_fps["cfg1_xyz"])
    assert np.array_equal(synth.lattice_clouds(3, 700, 4, 1), gold_fps["lattice_xyz"])
    assert np.array_equal(synth.duplicated_clouds(2, 1024, 0.05, 3), gold_fps["dups_xyz"])


@pytest.mark.parametrize("gen,m", [(lambda: synth.lattice_clouds(2, 1100, 3, 7), 200),
                                   (lambda: synth.duplicated_clouds(2, 600, 0.2, 8), 300),
                                   (lambda: synth.uniform_clouds(2, 50, 9), 60)])
def test_fps_closed_form_tie_rule_equals_literal_512_thread_simulation(gen, m):
    xyz = gen()
    assert np.array_equal(cref.fps(xyz, m), cref.fps(xyz, m, sim512=True))


def test_fps_properties():
    xyz = synth.uniform_clouds(3, 400, 4)
    idx = cref.fps(xyz, 400)
    assert (idx[:, 0] == 0).all()
    for b in range(3):                       # sampling all n points of a duplicate-free cloud is a permutation
        assert sorted(idx[b].tolist()) == list(range(400))
    # the min-distance-to-selected sequence is non-increasing
    p = xyz[0]
    sel = idx[0]
    d = np.full(400, np.inf)
    picked = []
    for j in range(1, 50):
        d = np.minimum(d, ((p - p[sel[j - 1]]) ** 2).sum(1))
        picked.append(d[sel[j]])
        assert np.isclose(d[sel[j]], d.max(), rtol=1e-5)
    assert all(a >= b - 1e-7 for a, b in zip(picked, picked[1:]))
    # thread count must not change results
    cref.set_threads(4)
    assert np.array_equal(cref.fps(xyz, 400), idx)
    cref.set_threads(1)


# ---- ball query / group / gather against the reference GPU kernels' recorded outputs
def test_ball_query_oracle_matches_reference_gpu_golden(gold_grp):
    xyz = gold_grp["cfg1_xyz"]
    assert np.array_equal(cref.fps(xyz, 512), gold_grp["cfg1_fps_idx"])
    q = cref.gather_point(xyz, gold_grp["cfg1_fps_idx"])
    assert np.array_equal(q, gold_grp["cfg1_gathered"])
    for tag, xyz2 in (("fpsq", q), ("self", xyz)):
        idx, cnt = cref.ball_query(0.2, 32, xyz, xyz2)
        assert np.array_equal(cnt, gold_grp[f"cfg1_{tag}_cnt"].astype(np.int32))
        assert np.array_equal(idx, gold_grp[f"cfg1_{tag}_idx"].astype(np.int32))
    lat = gold_grp["lat_xyz"]
    idx, cnt = cref.ball_query(1.0, 16, lat, lat[:, :300])
    assert np.array_equal(cnt, gold_grp["lat_cnt"].astype(np.int32))
    assert np.array_equal(idx, gold_grp["lat_idx"].astype(np.int32))


def test_ball_query_properties():
    xyz = synth.uniform_clouds(2, 300, 5)
    idx, cnt = cref.ball_query(0.3, 16, xyz, xyz)
    d = np.sqrt(((xyz[:, :, None] - xyz[:, None]) ** 2).sum(-1))
    for b in range(2):
        for j in range(0, 300, 17):
            inside = np.nonzero(d[b, j] < 0.3 - 1e-6)[0]
            got = idx[b, j, :cnt[b, j]]
            assert (np.diff(got) > 0).all()                        # ascending index order
            assert set(got.tolist()) <= set(np.nonzero(d[b, j] < 0.3 + 1e-6)[0].tolist())
            assert cnt[b, j] == min(16, len(np.nonzero(d[b, j] < 0.3)[0])) or abs(len(inside) - cnt[b, j]) <= 1
            assert (idx[b, j, cnt[b, j]:] == got[0]).all()         # padded with the first hit
            assert got[0] <= j                                     # a point is inside its own ball


def test_group_point_oracle_matches_reference_gpu_golden(gold_grp):
    out = cref.group_point(gold_grp["grp_pts"], gold_grp["grp_idx"])
    assert np.array_equal(out, gold_grp["grp_out"])
    g = cref.group_point_grad(256, gold_grp["grp_idx"], gold_grp["grp_grad_out"])
    np.testing.assert_allclose(g, gold_grp["grp_grad_points"], rtol=1e-5, atol=1e-6)   # atomics: order differs


def test_gather_grad_is_adjoint_of_gather():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 64, 3)).astype(np.float32)
    idx = rng.integers(0, 64, (2, 20)).astype(np.int32)
    g = rng.normal(size=(2, 20, 3)).astype(np.float32)
    lhs = (cref.gather_point(x, idx) * g).sum()
    rhs = (x * cref.gather_point_grad(64, idx, g)).sum()
    assert np.isclose(lhs, rhs, rtol=1e-5)


# ---- kNN / three_nn / three_interpolate against the reference's CPU code (golden + live when oracle/_ref exists)
def test_knn_oracle_matches_nanoflann_golden(gold_cpu):
    xyz = gold_cpu["knn_xyz"]
    assert np.array_equal(cref.knn(16, xyz, xyz[:, ::4]), gold_cpu["knn_idx_k16"].astype(np.int32))
    assert np.array_equal(cref.knn(32, xyz[:, :96], xyz[:, :96]), gold_cpu["knn_self_k32"].astype(np.int32))
    small = gold_cpu["knn_small_xyz"]
    assert np.array_equal(cref.knn(32, small, small), gold_cpu["knn_small_k32"].astype(np.int32))


def test_knn_properties():
    xyz = synth.uniform_clouds(2, 257, 12)
    d, idx = cref.knn(20, xyz, xyz, return_dist=True)
    assert (idx[:, :, 0] == np.arange(257)).all() and (d[:, :, 0] == 0).all()      # self first
    assert (np.diff(d, axis=-1) >= 0).all()                                        # ascending
    full = ((xyz[:, :, None] - xyz[:, None]) ** 2)
    full = (full[..., 0] + full[..., 1]) + full[..., 2]
    assert np.array_equal(np.sort(full, -1)[:, :, :20], d)                         # exactly the k smallest


def test_three_nn_and_interpolate_match_reference_cpu_golden(gold_cpu):
    xyz = gold_cpu["knn_xyz"]
    d, i = cref.three_nn(xyz, xyz[:, ::4])
    assert np.array_equal(d, gold_cpu["nn_dist"]) and np.array_equal(i, gold_cpu["nn_idx"].astype(np.int32))
    lat = gold_cpu["nn_lat_xyz"]
    d, i2 = cref.three_nn(lat, lat[:, :90])
    assert np.array_equal(d, gold_cpu["nn_lat_dist"]) and np.array_equal(i2, gold_cpu["nn_lat_idx"].astype(np.int32))
    d, i3 = cref.three_nn(xyz[:, :10], xyz[:, :2])
    assert np.array_equal(d, gold_cpu["nn_m2_dist"]) and np.array_equal(i3, gold_cpu["nn_m2_idx"].astype(np.int32))
    out = cref.three_interpolate(gold_cpu["ti_points"], i, gold_cpu["ti_weight"])
    assert np.array_equal(out, gold_cpu["ti_out"])                                 # bit-exact: same no-FMA arithmetic
    g = cref.three_interpolate_grad(128, i, gold_cpu["ti_weight"], gold_cpu["ti_grad_out"])
    assert np.array_equal(g, gold_cpu["ti_grad_points"])                           # same sequential accumulation order


@pytest.mark.skipif(not refnative.available("ref_knn.so"), reason="oracle/_ref not built (no /root/reference)")
def test_oracle_matches_compiled_reference_live():
    xyz = synth.uniform_clouds(4, 700, 21)
    q = xyz[:, :300]
    assert np.array_equal(cref.knn(32, xyz, q), refnative.knn_batch(xyz, q, 32).astype(np.int32))
    d, i = cref.three_nn(xyz, q)
    rd, ri = refnative.three_nn(xyz, q)
    assert np.array_equal(d, rd) and np.array_equal(i, ri)


def test_selection_sort_first_k_are_the_k_smallest():
    rng = np.random.default_rng(2)
    dist = rng.integers(0, 6, (2, 9, 40)).astype(np.float32)        # many ties
    outi, out = cref.selection_sort(7, dist)
    assert np.array_equal(out[:, :, :7], np.sort(dist, -1)[:, :, :7])
    assert np.array_equal(np.take_along_axis(dist, outi.astype(np.int64), 2), out)
    assert (np.sort(outi, -1) == np.arange(40)).all()               # a permutation of the row


# ---- grid subsampling: oracle == the reference's grid_subsampling.cpp (golden recorded through oracle/ref_shims)
def _sorted_rows(p, *rest):
    order = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
    return [p[order]] + [r[order] for r in rest]


@pytest.mark.parametrize("dl,tag", [(0.1, "dl10"), (0.04, "dl04"), (0.5, "dl50")])
def test_grid_subsampling_oracle_matches_reference_golden(gold_cpu, dl, tag):
    pts, feats, labels = gold_cpu["gs_points"], gold_cpu["gs_features"], gold_cpu["gs_labels"]
    p, f, lo = cref.grid_subsample(pts, feats, labels, dl)
    _, _, hi = cref.grid_subsample(pts, feats, labels, dl, tie_largest=True)
    order = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
    p, f, lo, hi = p[order], f[order], lo[order], hi[order]
    assert np.array_equal(p, gold_cpu[f"gs_{tag}_points"])          # bit-exact: fp32 sums in ascending point index
    assert np.array_equal(f, gold_cpu[f"gs_{tag}_features"])
    want = gold_cpu[f"gs_{tag}_labels"]
    unique_mode = lo == hi                                          # no count tie: the answer is well defined
    assert unique_mode.mean() > 0.5 and np.array_equal(lo[unique_mode], want[unique_mode])
    # under a tie the reference returns whatever std::unordered_map iterates first: one of the tied labels
    assert np.all((want >= lo) & (want <= hi))


def test_grid_subsampling_oracle_order_and_properties():
    rng = np.random.default_rng(4)
    pts = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    p = cref.grid_subsample(pts, dl=0.2)
    # cells come out in order of first appearance: the first cell holds point 0
    dl = np.float32(0.2)
    org = np.floor(pts.min(0) * (np.float32(1) / dl)) * dl                 # the reference's fp32 statements
    cell0 = np.floor((pts - org) / dl)
    same = (cell0 == cell0[0]).all(1)
    np.testing.assert_allclose(p[0], pts[same].mean(0), rtol=1e-5, atol=1e-6)
    assert len(p) == len(np.unique(cell0, axis=0))                         # one barycentre per occupied cell
    # a cell that contains the whole cloud gives one point: the centroid (the grid origin is a multiple of dl)
    pos = np.abs(pts) + np.float32(0.01)
    one = cref.grid_subsample(pos, dl=10.0)
    assert one.shape == (1, 3) and np.allclose(one[0], pos.mean(0), atol=1e-4)


@pytest.mark.skipif(not refnative.available("ref_gridsub.so"), reason="oracle/_ref not built (no /root/reference)")
def test_grid_subsampling_oracle_matches_compiled_reference_on_fresh_inputs():
    rng = np.random.default_rng(21)
    pts = (rng.normal(size=(20000, 3)) * np.array([3.0, 2.0, 0.5])).astype(np.float32)
    feats = rng.uniform(0, 1, (20000, 4)).astype(np.float32)
    for dl in (0.04, 0.3):
        a = _sorted_rows(*cref.grid_subsample(pts, feats, None, dl))
        b = _sorted_rows(*refnative.grid_subsampling(pts, feats, None, dl))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
