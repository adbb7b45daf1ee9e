"""Generates tests/golden/gpu_ref_*.npz: outputs of the REFERENCE's own GPU kernels (oracle/_ref/ref_sampling.so,
ref_grouping.so -- tf_sampling_g.cu / tf_grouping_g.cu compiled untouched by oracle/build_ref.sh) on seeded inputs.
Must run on a GPU box:  gpurun -- python tests/golden/make_golden_gpu.py   (writes gpurun_out/golden/, which is
then copied into tests/golden/ and committed).  The CPU oracle is pinned against these vectors in
tests/test_oracle_golden.py, which needs no GPU.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refnative  # noqa: E402
from tests import synth  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731

    # ---- FPS
    fps_cases = {
        "cfg1": (synth.uniform_clouds(2, 1024, 0), 512),
        "lattice": (synth.lattice_clouds(3, 700, 4, 1), 300),
        "dups": (synth.duplicated_clouds(2, 1024, 0.05, 3), 256),
        "n513": (synth.uniform_clouds(2, 513, 8), 200),
        "n3073": (synth.lattice_clouds(1, 3073, 6, 2), 200),
    }
    blob = {}
    for name, (xyz, m) in fps_cases.items():
        blob[f"{name}_xyz"] = xyz
        blob[f"{name}_idx"] = refnative.gpu_fps(t(xyz), m).cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "gpu_ref_fps.npz"), **blob)

    # ---- ball query + group_point + gather_point
    blob = {}
    xyz = synth.uniform_clouds(2, 1024, 0)
    fidx = refnative.gpu_fps(t(xyz), 512)
    q = refnative.gpu_gather_point(t(xyz), fidx).cpu().numpy()
    blob["cfg1_xyz"] = xyz
    blob["cfg1_fps_idx"] = fidx.cpu().numpy()
    blob["cfg1_gathered"] = q
    for tag, xyz2 in (("fpsq", q), ("self", xyz)):
        idx, cnt = refnative.gpu_ball_query(0.2, 32, t(xyz), t(xyz2))
        blob[f"cfg1_{tag}_idx"] = idx.cpu().numpy().astype(np.int16)
        blob[f"cfg1_{tag}_cnt"] = cnt.cpu().numpy().astype(np.int16)
    lat = synth.lattice_clouds(2, 900, 5, 3)
    idx, cnt = refnative.gpu_ball_query(1.0, 16, t(lat), t(lat[:, :300]))
    blob["lat_xyz"] = lat
    blob["lat_idx"] = idx.cpu().numpy().astype(np.int16)
    blob["lat_cnt"] = cnt.cpu().numpy().astype(np.int16)
    rng = np.random.default_rng(11)
    pts = rng.normal(size=(2, 256, 5)).astype(np.float32)
    gi = rng.integers(0, 256, (2, 32, 8)).astype(np.int32)
    blob["grp_pts"] = pts
    blob["grp_idx"] = gi
    blob["grp_out"] = refnative.gpu_group_point(t(pts), t(gi)).cpu().numpy()
    go = rng.normal(size=(2, 32, 8, 5)).astype(np.float32)
    blob["grp_grad_out"] = go
    blob["grp_grad_points"] = refnative.gpu_group_point_grad(256, t(gi), t(go)).cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "gpu_ref_grouping.npz"), **blob)
    print("golden written to", out_dir, os.listdir(out_dir))


if __name__ == "__main__":
    main()
