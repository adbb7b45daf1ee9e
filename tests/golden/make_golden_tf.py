"""Emits tests/golden/tf_graph_*.npz by EXECUTING THE REFERENCE'S OWN PYTHON (utils/pointasnl_util.py,
utils/pointnet_util.py, utils/tf_util.py, models/pointasnl_cls.py, models/pointasnl_sem_seg_res.py, imported
unmodified from /root/reference) on the stand-in tensorflow module tests/golden/tf_shim.py, in fp64.

Run here (needs /root/reference and oracle/_ref built):   python tests/golden/make_golden_tf.py

Each file holds the inputs (float32), the outputs (float64 results, stored as float32 where large) and, for every
variable and differentiable input, a gradient DIGEST of d(sum of outputs or loss)/d(variable): [sum, l2 norm,
projection on a name-seeded random direction] -- 3 numbers instead of the tensor, so the files stay small and no
weights need to be stored (variables are `oracle.golden_init.golden_init(name, shape, kind)` on every side).
tests/test_oracle.py pins oracle/model_ref.py to these files; tests/test_gpu_golden.py pins the CUDA path.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402
from oracle.golden_init import digest, probe_weights  # noqa: E402
from tests import synth  # noqa: E402


def grad_digests(y, named):
    """named: dict name -> tensor (requires_grad).  Returns {"grad/<name>": digest}."""
    names = list(named)
    grads = torch.autograd.grad(y, [named[n] for n in names], allow_unused=True)
    return {"grad/" + n: digest(n, np.zeros(tuple(named[n].shape)) if g is None else g.detach().numpy())
            for n, g in zip(names, grads)}


def trainables():
    return {k: v for k, v in tf_shim.variables().items() if v.requires_grad}


def moving(prefixes=None):
    return {"moving/" + k: v.detach().numpy().copy() for k, v in tf_shim.variables().items()
            if "moving_" in k}


def probe(t, name):
    return (t * torch.from_numpy(probe_weights(name, t.shape))).sum()


def save(name, blob):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(blob)} arrays")


# ------------------------------------------------------------------------------------------------ layer cases
def case_adaptive_sampling():
    tf_shim.install()
    import pointasnl_util as R
    rng = np.random.default_rng(11)
    gxyz = rng.uniform(-1, 1, (2, 24, 8, 3)).astype(np.float32)
    gfeat = np.concatenate([gxyz, rng.normal(size=(2, 24, 8, 10)).astype(np.float32)], -1)   # [xyz | feat] as grouping()
    xyz_t = tf_shim.tensor(gxyz).requires_grad_(True)
    feat_t = tf_shim.tensor(gfeat).requires_grad_(True)
    new_xyz, new_feature = R.AdaptiveSampling(xyz_t, feat_t, 6, True, 0.9, None, "as", True)
    y = probe(new_xyz, "as/new_xyz") + probe(new_feature, "as/new_feature")
    named = dict(trainables(), **{"in/group_xyz": xyz_t, "in/group_feature": feat_t})
    blob = {"group_xyz": gxyz, "group_feature": gfeat, "new_xyz": new_xyz.detach().numpy(),
            "new_feature": new_feature.detach().numpy(), **grad_digests(y, named), **moving()}
    save("tf_graph_adaptive_sampling.npz", blob)


def case_nonlocal():
    tf_shim.install()
    import pointasnl_util as R
    rng = np.random.default_rng(12)
    feature = rng.normal(size=(2, 96, 20)).astype(np.float32)
    new_point = rng.normal(size=(2, 1, 40, 23)).astype(np.float32)
    f_t = tf_shim.tensor(feature).requires_grad_(True)
    p_t = tf_shim.tensor(new_point).requires_grad_(True)
    out = R.PointNonLocalCell(f_t, p_t, [32, 48], True, 0.9, None, "nl", True)
    y = probe(out, "nl/out")
    named = dict(trainables(), **{"in/feature": f_t, "in/new_point": p_t})
    save("tf_graph_nonlocal.npz", {"feature": feature, "new_point": new_point, "out": out.detach().numpy(),
                                   **grad_digests(y, named), **moving()})


def case_nonlocal_concat():
    tf_shim.install()
    import pointasnl_util as R
    rng = np.random.default_rng(18)
    feature = rng.normal(size=(2, 48, 12)).astype(np.float32)
    new_point = rng.normal(size=(2, 1, 20, 15)).astype(np.float32)
    f_t = tf_shim.tensor(feature).requires_grad_(True)
    p_t = tf_shim.tensor(new_point).requires_grad_(True)
    out = R.PointNonLocalCell(f_t, p_t, [32, 24], True, 0.9, None, "nlc", True, mode="concat")
    y = probe(out, "nlc/out")
    named = dict(trainables(), **{"in/feature": f_t, "in/new_point": p_t})
    save("tf_graph_nonlocal_concat.npz", {"feature": feature, "new_point": new_point, "out": out.detach().numpy(),
                                          **grad_digests(y, named), **moving()})


def case_set_abstraction():
    blob = {}
    rng = np.random.default_rng(13)
    xyz = synth.uniform_clouds(2, 160, 13)
    feat = rng.normal(size=(2, 160, 8)).astype(np.float32)
    blob["xyz"], blob["feature"] = xyz, feat
    # (tag, npoint, nsample, mlp, as_neighbor, NL)
    for tag, npoint, nsample, mlp, as_nb, nl in (("as0", 40, 16, [16, 16, 32], 0, True),
                                                 ("as6", 40, 16, [16, 16, 32], 6, True),
                                                 ("self", 160, 16, [32, 32], 0, False)):
        tf_shim.install()
        import pointasnl_util as R
        x_t = tf_shim.tensor(xyz).requires_grad_(True)
        f_t = tf_shim.tensor(feat).requires_grad_(True)
        new_xyz, new_point = R.PointASNLSetAbstraction(x_t, f_t, npoint, nsample, mlp, True, 0.9, None, "layer1",
                                                       as_neighbor=as_nb, NL=nl)
        y = probe(new_point, tag + "/new_point") + (probe(new_xyz, tag + "/new_xyz") if as_nb > 0 else 0.0)
        named = dict(trainables(), **{"in/xyz": x_t, "in/feature": f_t})
        blob[tag + "/new_xyz"] = new_xyz.detach().numpy()
        blob[tag + "/new_point"] = new_point.detach().numpy()
        blob.update({tag + "/" + k: v for k, v in grad_digests(y, named).items()})
        blob.update({tag + "/" + k: v for k, v in moving().items()})
    save("tf_graph_set_abstraction.npz", blob)


def case_fp_module():
    tf_shim.install()
    import pointnet_util as P
    rng = np.random.default_rng(14)
    xyz1 = synth.uniform_clouds(2, 120, 14)
    xyz2 = xyz1[:, ::4].copy() + 0.0
    points1 = rng.normal(size=(2, 120, 12)).astype(np.float32)
    points2 = rng.normal(size=(2, 30, 20)).astype(np.float32)
    p1, p2 = tf_shim.tensor(points1).requires_grad_(True), tf_shim.tensor(points2).requires_grad_(True)
    out = P.pointnet_fp_module(tf_shim.tensor(xyz1), tf_shim.tensor(xyz2), p1, p2, [32, 24], True, 0.9, "fa", bn=True)
    y = probe(out, "fa/out")
    named = dict(trainables(), **{"in/points1": p1, "in/points2": p2})
    save("tf_graph_fp_module.npz", {"xyz1": xyz1, "xyz2": xyz2, "points1": points1, "points2": points2,
                                    "out": out.detach().numpy(), **grad_digests(y, named), **moving()})


def case_decoding_layer():
    tf_shim.install()
    import pointasnl_util as R
    rng = np.random.default_rng(15)
    xyz1 = synth.uniform_clouds(2, 96, 15)
    xyz2 = xyz1[:, ::4].copy()
    points1 = rng.normal(size=(2, 96, 12)).astype(np.float32)
    points2 = rng.normal(size=(2, 24, 16)).astype(np.float32)
    p1, p2 = tf_shim.tensor(points1).requires_grad_(True), tf_shim.tensor(points2).requires_grad_(True)
    out = R.PointASNLDecodingLayer(tf_shim.tensor(xyz1), tf_shim.tensor(xyz2), p1, p2, 8, [32, 24], True, 0.9, None,
                                   "up", NL=False)
    y = probe(out, "up/out")
    named = dict(trainables(), **{"in/points1": p1, "in/points2": p2})
    save("tf_graph_decoding_layer.npz", {"xyz1": xyz1, "xyz2": xyz2, "points1": points1, "points2": points2,
                                         "out": out.detach().numpy(), **grad_digests(y, named), **moving()})


def case_repulsion():
    tf_shim.install()
    import pointasnl_util as R
    blob = {}
    for tag, n, radius, seed in (("dense", 256, 0.3, 16), ("sparse", 200, 0.07, 17)):
        pred = synth.uniform_clouds(2, n, seed)
        if tag == "sparse":
            pred *= 0.5                              # several points per ball, some balls with < 5 members
        p = tf_shim.tensor(pred).requires_grad_(True)
        loss = R.get_repulsion_loss(p, nsample=20, radius=radius)
        (g,) = torch.autograd.grad(loss, [p])
        blob[tag + "/pred"], blob[tag + "/radius"] = pred, np.float64(radius)
        blob[tag + "/loss"], blob[tag + "/grad"] = loss.detach().numpy(), g.numpy()
    save("tf_graph_repulsion.npz", blob)


# ------------------------------------------------------------------------------------------------ models
def case_cls():
    blob = {}
    pc, label = synth.modelnet_batch(4, 640, 21)
    blob["point_cloud"], blob["label"] = pc, label.astype(np.int32)
    for tag, adaptive, uw in (("plain", False, 0.0), ("as", True, 0.0), ("as_uniform", True, 0.5)):
        tf = tf_shim.install()
        import pointasnl_cls as M
        t0 = time.time()
        pred, end_points = M.get_model(tf_shim.tensor(pc), tf.constant(True), use_normal=True, bn_decay=0.9,
                                       adaptive_sample=adaptive)
        loss = M.get_loss(pred, tf_shim.tensor(label), end_points, uniform_weight=uw)
        blob[tag + "/pred"] = pred.detach().numpy()
        blob[tag + "/loss"] = loss.detach().numpy()
        blob[tag + "/l1_xyz"] = end_points["l1_xyz"].detach().numpy()
        blob.update({tag + "/" + k: v for k, v in grad_digests(loss, trainables()).items()})
        if tag == "plain":
            blob.update({tag + "/" + k: v for k, v in moving().items() if "fc1" in k or "layer1/conv0" in k
                         or "layer2/aggregation" in k})
        print(f"cls {tag}: loss {float(loss):.6f}  {time.time() - t0:.1f}s")
    save("tf_graph_cls.npz", blob)


def case_seg():
    blob = {}
    stride = 16
    for tag, fc, seed in (("scannet", 3, 22), ("kitti", 1, 23)):
        tf = tf_shim.install()
        import pointasnl_sem_seg_res as M
        if tag == "scannet":
            pts, label, smpw = synth.scannet_batch(2, 8192, seed)
            ncls, radius = 21, 0.07
        else:
            pts, label = synth.kitti_batch(2, 8192, seed)
            smpw = np.random.default_rng(seed).uniform(1.0, 4.9, label.shape).astype(np.float32)
            ncls, radius = 20, 0.2
        smpw = smpw.copy()
        smpw[:, ::7] = 0.0                          # unlabelled points: weight 0 (SUM_BY_NONZERO_WEIGHTS matters)
        t0 = time.time()
        pred, end_points = M.get_model(tf_shim.tensor(pts), tf.constant(True), ncls, bn_decay=0.9, feature_channel=fc)
        loss = M.get_loss(pred, tf_shim.tensor(label), end_points, smpw=tf_shim.tensor(smpw), radius=radius)
        blob[tag + "/points"], blob[tag + "/label"], blob[tag + "/smpw"] = pts, label.astype(np.int32), smpw
        blob[tag + "/pred_strided"] = pred.detach().numpy()[:, ::stride].astype(np.float32)
        blob[tag + "/pred_digest"] = digest(tag + "/pred", pred.detach().numpy())
        blob[tag + "/loss"] = loss.detach().numpy()
        blob[tag + "/l1_xyz"] = end_points["l1_xyz"].detach().numpy().astype(np.float32)
        blob.update({tag + "/" + k: v for k, v in grad_digests(loss, trainables()).items()})
        print(f"seg {tag}: loss {float(loss):.6f}  {time.time() - t0:.1f}s")
    save("tf_graph_seg.npz", blob)


CASES = {"as": case_adaptive_sampling, "nl": case_nonlocal, "nlc": case_nonlocal_concat, "sa": case_set_abstraction, "fp": case_fp_module,
         "dec": case_decoding_layer, "rep": case_repulsion, "cls": case_cls, "seg": case_seg}

if __name__ == "__main__":
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
