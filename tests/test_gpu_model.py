"""Model-level parity on the GPU box: the drop-in callers (pointasnl_b200.callers) driving the CUDA kernels vs the
line-by-line CPU restatement of the reference's TF graph code (oracle/model_ref.py) with THE SAME variables.

Tolerances: the hot-path kernels are held to 1e-4 relative in tests/test_gpu_parity.py.  A whole network stacks
~40 batch-normalised fp32 layers on top, and batch norm divides by small standard deviations, so model-level
outputs are compared against an fp64 evaluation of the oracle with the tolerance measured from the oracle's own
fp32-vs-fp64 gap (printed by the tests): logits within 2e-3, loss within 1e-4 relative, every parameter gradient
within 2 % of its tensor norm.  Discrete decisions (FPS picks, neighbour indices) depend on fp32 coordinates only
and are bit-exact, which is what keeps this comparison meaningful.
"""
import contextlib

import numpy as np
import pytest
import torch

from oracle import model_ref
from oracle.golden_init import golden_init
from tests import synth


def named_init(full, shape, leaf):
    """name-seeded variable values: independent of the order in which the layers create their variables"""
    return golden_init(full, shape, leaf)

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def no_dropout():
    """dropout draws from different RNG streams on the two sides; parity runs compare the deterministic part."""
    from pointasnl_b200 import tf_util
    orig = tf_util.dropout
    tf_util.dropout = lambda inputs, is_training, scope, keep_prob=0.5, noise_shape=None: inputs
    try:
        yield
    finally:
        tf_util.dropout = orig


def rel_l2(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def check_param_grads(grads, ref):
    """Every parameter gradient within 2 % of its tensor norm.  Some gradients are identically zero in exact
    arithmetic (e.g. the BN beta of the attention KEY projection: softmax is invariant to a per-query constant added
    to all logits), where only rounding noise remains; those are held to an absolute floor of 1e-5 of the largest
    gradient norm in the model instead of a meaningless relative bar."""
    assert set(grads) == set(ref)
    scale = max(float(ref[n].double().norm()) for n in ref)
    rows = []
    for n in grads:
        err = float((grads[n].double() - ref[n].double()).norm())
        bar = 2e-2 * float(ref[n].double().norm()) + 1e-5 * scale
        rows.append((err / bar, n, err, float(ref[n].double().norm())))
    rows.sort(reverse=True)
    print("largest gradient errors relative to their bar (err/bar, name, |err|, |ref|):")
    for r in rows[:5]:
        print("   %.3f  %-45s %.3e %.3e" % r)
    assert rows[0][0] < 1.0, rows[0]


def test_sa_layer_with_adaptive_sampling_matches_oracle(cuda):
    """cls layer1 with --AS: FPS 1024->512, kNN 32, AS over 12 neighbours, NL cell, local cell; forward + backward."""
    from pointasnl_b200 import tf_util
    from pointasnl_b200.pointasnl_util import PointASNLSetAbstraction
    pc, _ = synth.modelnet_batch(4, 1024, 1)
    xyz = torch.from_numpy(pc[:, :, :3]).to(cuda)
    feat = torch.from_numpy(pc[:, :, 3:]).to(cuda).requires_grad_(True)
    store = tf_util.VariableStore(cuda, seed=3, init_fn=named_init)
    with store.as_default():
        new_xyz, new_pts = PointASNLSetAbstraction(xyz, feat, 512, 32, [64, 64, 128], True, None, None, "layer1",
                                                   as_neighbor=12)
    g1, g2 = torch.randn_like(new_xyz), torch.randn_like(new_pts)
    ((new_xyz * g1).sum() + (new_pts * g2).sum()).backward()
    grads = {n: v.grad.detach().cpu() for n, v in store.variables().items() if v.requires_grad}

    res = {}
    for dt in (torch.float32, torch.float64):
        r = model_ref.Ref(store.variables(), dtype=dt)
        f = torch.from_numpy(pc[:, :, 3:]).to(dt).requires_grad_(True)
        ox, op = model_ref.set_abstraction(r, torch.from_numpy(pc[:, :, :3]).to(dt), f, 512, 32, [64, 64, 128], "layer1",
                                           as_neighbor=12)
        ((ox * g1.cpu().to(dt)).sum() + (op * g2.cpu().to(dt)).sum()).backward()
        res[dt] = (ox.detach(), op.detach(), {n: v.grad for n, v in r.v.items() if v.grad is not None}, f.grad)
    ox, op, og, of = res[torch.float64]
    print("oracle fp32-vs-fp64 gap: xyz %.2e feat %.2e" % (rel_l2(res[torch.float32][0], ox), rel_l2(res[torch.float32][1], op)))
    print("cuda-vs-fp64:            xyz %.2e feat %.2e" % (rel_l2(new_xyz.cpu(), ox), rel_l2(new_pts.cpu(), op)))
    torch.testing.assert_close(new_xyz.detach().cpu().double(), ox, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(new_pts.detach().cpu().double(), op, rtol=2e-3, atol=2e-3)
    assert rel_l2(feat.grad.cpu(), of) < 2e-2
    check_param_grads(grads, og)


@pytest.mark.parametrize("adaptive_sample", [False, True])
def test_cls_model_forward_backward_matches_oracle(cuda, adaptive_sample):
    """BASELINE config 2 at reduced batch: pointasnl_cls fwd + loss + bwd on (8, 1024, 6)."""
    from pointasnl_b200 import tf_util
    from pointasnl_b200.callers import pointasnl_cls
    pc, labels = synth.modelnet_batch(8, 1024, 1)
    x = torch.from_numpy(pc).to(cuda)
    y = torch.from_numpy(labels).to(cuda)
    # --AS: layer 2 samples and searches on COMPUTED coordinates (the Adaptive-Sampling output), where a last-bit
    # difference between the fp32 device run and the fp64 oracle can flip a near-tie in FPS / k-NN and change the logits
    # at the 1e-3 level; the TF initialisers (gamma 1, beta 0) at this seed produce no such flip, name-seeded values do.
    # (tests/test_gpu_golden.py covers --AS with non-trivial variables on an input without near-ties.)
    store = tf_util.VariableStore(cuda, seed=1, init_fn=None if adaptive_sample else named_init)
    with store.as_default(), no_dropout():
        pred, end_points = pointasnl_cls.get_model(x, True, use_normal=True, adaptive_sample=adaptive_sample)
        loss = pointasnl_cls.get_loss(pred, y, end_points)
    loss.backward()
    grads = {n: v.grad.detach().cpu() for n, v in store.variables().items() if v.requires_grad}
    assert all(torch.isfinite(g).all() for g in grads.values())

    out = {}
    for dt in (torch.float32, torch.float64):
        r = model_ref.Ref(store.variables(), dtype=dt, dropout=False)
        net, total = model_ref.cls_forward(r, torch.from_numpy(pc), torch.from_numpy(labels), use_normal=True,
                                           adaptive_sample=adaptive_sample)
        total.backward()
        out[dt] = (net.detach(), total.detach(), {n: v.grad for n, v in r.v.items() if v.grad is not None})
    net64, loss64, g64 = out[torch.float64]
    print("oracle fp32-vs-fp64 logits gap %.2e; cuda-vs-fp64 %.2e" % (rel_l2(out[torch.float32][0], net64), rel_l2(pred.cpu(), net64)))
    torch.testing.assert_close(pred.detach().cpu().double(), net64, rtol=2e-3, atol=2e-3)
    assert abs(float(loss) - float(loss64)) <= 1e-4 * abs(float(loss64)) + 1e-5
    check_param_grads(grads, g64)


def test_seg_res_model_forward_backward_runs_and_matches_first_stage(cuda):
    """BASELINE config 4 at B=1, N=8192 (the smallest N for which the deepest layer still has >= 32 points)."""
    from pointasnl_b200 import tf_util
    from pointasnl_b200.callers import pointasnl_sem_seg_res as seg
    pts, labels, smpw = synth.scannet_batch(1, 8192, 2)
    x = torch.from_numpy(pts).to(cuda)
    store = tf_util.VariableStore(cuda, seed=2, init_fn=named_init)
    with store.as_default(), no_dropout():
        pred, end_points = seg.get_model(x, True, 21, feature_channel=3)
        loss = seg.get_loss(pred, torch.from_numpy(labels).to(cuda), end_points, torch.from_numpy(smpw).to(cuda))
    loss.backward()
    assert tuple(pred.shape) == (1, 8192, 21) and torch.isfinite(pred).all() and torch.isfinite(loss)
    assert all(torch.isfinite(v.grad).all() for v in store.parameters())
    # Everything up to and including layer1_1 depends on raw input coordinates only, so its discrete decisions are
    # exact and l1_xyz (the Adaptive-Sampling output) must match the oracle.  Later layers search neighbours on
    # l1_xyz itself; a 1-ulp difference there may legitimately flip a near-tie, so they are covered by the
    # per-op parity tests instead of a brittle end-to-end comparison.
    r = model_ref.Ref(store.variables(), dtype=torch.float64, dropout=False)
    p = torch.from_numpy(pts).double()
    _, l0 = model_ref.set_abstraction(r, p[:, :, :3], p[:, :, 3:], 8192, 32, [16, 16, 32], "layer0", as_neighbor=0, NL=False)
    l1_xyz, _ = model_ref.set_abstraction(r, p[:, :, :3], l0, 1024, 32, [32, 32, 64], "layer1_1", as_neighbor=8)
    torch.testing.assert_close(end_points["l1_xyz"].detach().cpu().double(), l1_xyz, rtol=1e-3, atol=1e-4)
