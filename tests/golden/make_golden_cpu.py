"""Generates tests/golden/cpu_ref.npz from the REFERENCE's own CPU code compiled untouched into oracle/_ref by
oracle/build_ref.sh: nanoflann kNN (utils/nearest_neighbors/knn_.cxx) and threenn_cpu / threeinterpolate_cpu /
threeinterpolate_grad_cpu (tf_ops/3d_interpolation/tf_interpolate.cpp:60-153).  Run here (needs /root/reference
to have been compiled):  python tests/golden/make_golden_cpu.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refnative  # noqa: E402
from tests import synth  # noqa: E402


def main():
    blob = {}
    xyz = synth.uniform_clouds(2, 512, 0)
    q = xyz[:, ::4].copy()
    blob["knn_xyz"] = xyz
    blob["knn_idx_k16"] = refnative.knn_batch(xyz, q, 16).astype(np.int16)
    blob["knn_self_k32"] = refnative.knn_batch(xyz[:, :96], xyz[:, :96], 32).astype(np.int16)
    small = synth.uniform_clouds(3, 40, 6)
    blob["knn_small_xyz"] = small
    blob["knn_small_k32"] = refnative.knn_batch(small, small, 32).astype(np.int16)   # K = 80 % of the cloud
    # three_nn on random and on lattice (ties) inputs
    d, i = refnative.three_nn(xyz, q)
    blob["nn_dist"], blob["nn_idx"] = d, i.astype(np.int16)
    lat = synth.lattice_clouds(2, 300, 4, 5)
    d2, i2 = refnative.three_nn(lat, lat[:, :90])
    blob["nn_lat_xyz"], blob["nn_lat_dist"], blob["nn_lat_idx"] = lat, d2, i2.astype(np.int16)
    d3, i3 = refnative.three_nn(xyz[:, :10], xyz[:, :2])     # fewer than 3 known points
    blob["nn_m2_dist"], blob["nn_m2_idx"] = d3, i3.astype(np.int16)
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(2, 128, 7)).astype(np.float32)
    w = rng.uniform(0.1, 1.0, (2, 512, 3)).astype(np.float32)
    blob["ti_points"], blob["ti_weight"] = pts, w
    blob["ti_out"] = refnative.three_interpolate(pts, i, w)
    g = rng.normal(size=(2, 512, 7)).astype(np.float32)
    blob["ti_grad_out"] = g
    blob["ti_grad_points"] = refnative.three_interpolate_grad(128, i, w, g)
    # grid subsampling: the reference's grid_subsampling.cpp (through oracle/ref_shims), rows sorted lexicographically
    # because the reference's own output order is std::unordered_map iteration order
    def sorted_rows(p, *rest):
        order = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
        return [p[order]] + [r[order] for r in rest]
    rng = np.random.default_rng(9)
    gp = (rng.uniform(-1.0, 1.0, (3000, 3)) * np.array([2.0, 1.5, 0.7])).astype(np.float32)
    gp[100:200] = gp[0:100]                                  # exact duplicates
    gf = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    gl = np.repeat(rng.integers(0, 5, (750, 2)), 4, axis=0).astype(np.int32)     # runs of equal labels: few count ties
    blob["gs_points"], blob["gs_features"], blob["gs_labels"] = gp, gf, gl
    for dl, tag in ((0.1, "dl10"), (0.04, "dl04"), (0.5, "dl50")):
        p_, f_, l_ = refnative.grid_subsampling(gp, gf, gl, dl)
        p_, f_, l_ = sorted_rows(p_, f_, l_)
        blob[f"gs_{tag}_points"], blob[f"gs_{tag}_features"], blob[f"gs_{tag}_labels"] = p_, f_, l_
    out = os.path.join(ROOT, "tests", "golden", "cpu_ref.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
