"""Pins oracle/model_ref.py (the PyTorch-CPU restatement of the reference's TF graph code) to
tests/golden/tf_graph_*.npz -- vectors produced by executing the reference's OWN Python files on the stand-in
tensorflow module (tests/golden/tf_shim.py, generator tests/golden/make_golden_tf.py).  fp64 on both sides, the same
name-seeded variables (oracle/golden_init.py): outputs must agree to 1e-9 relative, the gradient digest of every
variable and differentiable input to 1e-7.  No GPU, no /root/reference needed (the vectors are committed)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref as M
from oracle.golden_init import digest, probe_weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64 = torch.float64


def load(name):
    return np.load(os.path.join(GOLD, name))


def t(a):
    return torch.from_numpy(np.asarray(a)).to(F64)


def probe(x, name):
    return (x * torch.from_numpy(probe_weights(name, x.shape))).sum()


def close(got, want, rtol=1e-9, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    assert err <= rtol, f"{what}: max error {err:.3e} of the largest value (bar {rtol:g})"


def check_grads(gold, prefix, y, named, rtol=1e-7, atol=1e-11):
    """every `grad/<name>` digest in the golden file must be matched: no variable may be missing on either side."""
    keys = [k for k in gold.files if k.startswith(prefix + "grad/")]
    want_names = {k[len(prefix) + 5:] for k in keys}
    assert want_names == set(named), (sorted(want_names ^ set(named)))
    names = sorted(named)
    grads = torch.autograd.grad(y, [named[n] for n in names], allow_unused=True)
    for n, g in zip(names, grads):
        want = gold[prefix + "grad/" + n]
        got = digest(n, np.zeros(tuple(named[n].shape)) if g is None else g.detach().numpy())
        scale = want[1]                              # the l2 norm of the reference gradient
        # atol: gradients that are analytically zero (a bias in front of a train-mode batch norm) are rounding noise
        assert np.abs(got - want).max() <= (rtol * scale + atol) * max(1.0, np.sqrt(named[n].numel())), (n, got, want)


def ref():
    return M.Ref(None, dtype=F64, dropout=False, init="golden")


def trainables(r):
    return {k: v for k, v in r.v.items()}


def test_adaptive_sampling_matches_reference_python():
    g = load("tf_graph_adaptive_sampling.npz")
    r = ref()
    xyz, feat = t(g["group_xyz"]).requires_grad_(True), t(g["group_feature"]).requires_grad_(True)
    new_xyz, new_feature = M.adaptive_sampling(r, xyz, feat, 6, "as", True)
    close(new_xyz.detach(), g["new_xyz"], what="new_xyz")
    close(new_feature.detach(), g["new_feature"], what="new_feature")
    y = probe(new_xyz, "as/new_xyz") + probe(new_feature, "as/new_feature")
    check_grads(g, "", y, dict(trainables(r), **{"in/group_xyz": xyz, "in/group_feature": feat}))


def test_nonlocal_cell_matches_reference_python():
    g = load("tf_graph_nonlocal.npz")
    r = ref()
    f, p = t(g["feature"]).requires_grad_(True), t(g["new_point"]).requires_grad_(True)
    out = M.point_nonlocal_cell(r, f, p, [32, 48], "nl", True)
    close(out.detach(), g["out"], what="out")
    check_grads(g, "", probe(out, "nl/out"), dict(trainables(r), **{"in/feature": f, "in/new_point": p}))


@pytest.mark.parametrize("tag,npoint,nsample,mlp,as_nb,nl", [("as0", 40, 16, [16, 16, 32], 0, True),
                                                              ("as6", 40, 16, [16, 16, 32], 6, True),
                                                              ("self", 160, 16, [32, 32], 0, False)])
def test_set_abstraction_matches_reference_python(tag, npoint, nsample, mlp, as_nb, nl):
    g = load("tf_graph_set_abstraction.npz")
    r = ref()
    xyz, feat = t(g["xyz"]).requires_grad_(True), t(g["feature"]).requires_grad_(True)
    new_xyz, new_point = M.set_abstraction(r, xyz, feat, npoint, nsample, mlp, "layer1", as_neighbor=as_nb, NL=nl)
    close(new_xyz.detach(), g[tag + "/new_xyz"], what="new_xyz")
    close(new_point.detach(), g[tag + "/new_point"], what="new_point")
    y = probe(new_point, tag + "/new_point") + (probe(new_xyz, tag + "/new_xyz") if as_nb > 0 else 0.0)
    check_grads(g, tag + "/", y, dict(trainables(r), **{"in/xyz": xyz, "in/feature": feat}))


def test_fp_module_matches_reference_python():
    g = load("tf_graph_fp_module.npz")
    r = ref()
    p1, p2 = t(g["points1"]).requires_grad_(True), t(g["points2"]).requires_grad_(True)
    out = M.fp_module(r, t(g["xyz1"]), t(g["xyz2"]), p1, p2, [32, 24], "fa")
    # the reference's three_interpolate is an fp32 C++ op: the golden side rounded its operands to fp32
    close(out.detach(), g["out"], rtol=2e-6, what="out")
    check_grads(g, "", probe(out, "fa/out"), dict(trainables(r), **{"in/points1": p1, "in/points2": p2}), rtol=2e-6)


def test_decoding_layer_matches_reference_python():
    g = load("tf_graph_decoding_layer.npz")
    r = ref()
    p1, p2 = t(g["points1"]).requires_grad_(True), t(g["points2"]).requires_grad_(True)
    out = M.decoding_layer(r, t(g["xyz1"]), t(g["xyz2"]), p1, p2, 8, [32, 24], "up")
    close(out.detach(), g["out"], rtol=2e-6, what="out")
    check_grads(g, "", probe(out, "up/out"), dict(trainables(r), **{"in/points1": p1, "in/points2": p2}), rtol=2e-6)


@pytest.mark.parametrize("tag", ["dense", "sparse"])
def test_repulsion_loss_matches_reference_python(tag):
    g = load("tf_graph_repulsion.npz")
    pred = t(g[tag + "/pred"]).requires_grad_(True)
    loss = M.get_repulsion_loss(None, pred, 20, float(g[tag + "/radius"]))
    # group_point is an fp32 op in the reference; pred holds fp32 values, so nothing is lost
    close(loss.detach(), g[tag + "/loss"], what="loss")
    (grad,) = torch.autograd.grad(loss, [pred])
    close(grad, g[tag + "/grad"], rtol=1e-7, what="grad")


@pytest.mark.parametrize("tag,adaptive,uw", [("plain", False, 0.0), ("as", True, 0.0), ("as_uniform", True, 0.5)])
def test_cls_model_matches_reference_python(tag, adaptive, uw):
    g = load("tf_graph_cls.npz")
    r = ref()
    pc, label = t(g["point_cloud"]), torch.from_numpy(g["label"]).long()
    pred, loss = M.cls_forward(r, pc, label, use_normal=True, adaptive_sample=adaptive, uniform_weight=uw)
    close(pred.detach(), g[tag + "/pred"], rtol=1e-8, what="pred")
    close(loss.detach(), g[tag + "/loss"], rtol=1e-9, what="loss")
    check_grads(g, tag + "/", loss, trainables(r), rtol=1e-6)


@pytest.mark.parametrize("tag,fc,ncls,radius", [("scannet", 3, 21, 0.07), ("kitti", 1, 20, 0.2)])
def test_seg_model_matches_reference_python(tag, fc, ncls, radius):
    g = load("tf_graph_seg.npz")
    r = ref()
    pts, label, smpw = t(g[tag + "/points"]), torch.from_numpy(g[tag + "/label"]).long(), t(g[tag + "/smpw"])
    pred, loss = M.seg_forward(r, pts, ncls, label, smpw, feature_channel=fc, radius=radius)
    close(pred.detach()[:, ::16], g[tag + "/pred_strided"], rtol=2e-6, what="pred")   # stored as float32
    got = digest(tag + "/pred", pred.detach().numpy())
    assert np.abs(got - g[tag + "/pred_digest"]).max() <= 1e-5 * g[tag + "/pred_digest"][1]
    close(loss.detach(), g[tag + "/loss"], rtol=1e-6, what="loss")
    check_grads(g, tag + "/", loss, trainables(r), rtol=2e-5)
