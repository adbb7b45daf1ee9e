"""GPU grid subsampling (csrc/grid_subsampling.cu, pointasnl_b200.grid_subsampling) against the C oracle and the
golden vectors recorded from the reference's grid_subsampling.cpp: barycentres and feature means bit-exact, cells in
order of first appearance, label ties to the smallest label."""
import os

import numpy as np
import pytest
import torch

from oracle import cref, refnative

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_matches_reference_golden(cuda):
    from pointasnl_b200.grid_subsampling import grid_subsampling
    g = np.load(os.path.join(GOLD, "cpu_ref.npz"))
    pts, feats, labels = g["gs_points"], g["gs_features"], g["gs_labels"]
    for dl, tag in ((0.1, "dl10"), (0.04, "dl04"), (0.5, "dl50")):
        p, f, lab = grid_subsampling(pts, features=feats, labels=labels, sampleDl=dl)
        op, of, ol = cref.grid_subsample(pts, feats, labels, dl)
        assert np.array_equal(p, op) and np.array_equal(f, of) and np.array_equal(lab, ol)    # same order, same bits
        order = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
        assert np.array_equal(p[order], g[f"gs_{tag}_points"]) and np.array_equal(f[order], g[f"gs_{tag}_features"])


@pytest.mark.parametrize("n,dl", [(1, 0.1), (7, 0.5), (100000, 0.04), (1000000, 0.06)])
def test_matches_oracle_on_scannet_like_and_edge_inputs(cuda, n, dl):
    from pointasnl_b200.grid_subsampling import grid_subsampling
    rng = np.random.default_rng(n)
    pts = (rng.normal(size=(n, 3)) * np.array([2.5, 2.0, 0.6])).astype(np.float32)
    rgb = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    lab = rng.integers(0, 21, n).astype(np.int32)
    p, f, l1 = grid_subsampling(torch.from_numpy(pts).to(cuda), features=torch.from_numpy(rgb).to(cuda),
                                labels=torch.from_numpy(lab).to(cuda), sampleDl=dl)
    op, of, ol = cref.grid_subsample(pts, rgb, lab, dl)
    assert tuple(p.shape) == op.shape and l1.dim() == 1
    assert np.array_equal(p.cpu().numpy(), op) and np.array_equal(f.cpu().numpy(), of)
    assert np.array_equal(l1.cpu().numpy(), ol[:, 0])
    only = grid_subsampling(torch.from_numpy(pts).to(cuda), sampleDl=dl)          # points only: a single tensor back
    assert torch.is_tensor(only) and np.array_equal(only.cpu().numpy(), op)


@pytest.mark.skipif(not refnative.available("ref_gridsub.so"), reason="oracle/_ref not built")
def test_matches_compiled_reference_as_a_set(cuda):
    from pointasnl_b200.grid_subsampling import grid_subsampling
    rng = np.random.default_rng(77)
    pts = rng.uniform(-2, 2, (50000, 3)).astype(np.float32)
    pts[1000:3000] = pts[:2000]
    feats = rng.normal(size=(50000, 2)).astype(np.float32)
    p, f = grid_subsampling(pts, features=feats, sampleDl=0.1)
    rp, rf = refnative.grid_subsampling(pts, feats, None, 0.1)
    o1, o2 = np.lexsort((p[:, 2], p[:, 1], p[:, 0])), np.lexsort((rp[:, 2], rp[:, 1], rp[:, 0]))
    assert np.array_equal(p[o1], rp[o2]) and np.array_equal(f[o1], rf[o2])
