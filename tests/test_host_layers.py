"""Host-side (CPU, op-by-op path) checks of layer code that has no CUDA kernel of its own, against the vectors produced
by the reference's own Python (tests/golden/tf_graph_*.npz)."""
import os

import numpy as np
import torch

from oracle.golden_init import digest, golden_init, probe_weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_nonlocal_cell_concat_mode_matches_reference_python_on_cpu():
    from pointasnl_b200 import tf_util
    from pointasnl_b200.pointasnl_util import PointNonLocalCell
    g = np.load(os.path.join(GOLD, "tf_graph_nonlocal_concat.npz"))
    store = tf_util.VariableStore("cpu", init_fn=lambda full, shape, leaf: golden_init(full, shape, leaf))
    f = torch.from_numpy(g["feature"]).requires_grad_(True)
    p = torch.from_numpy(g["new_point"]).requires_grad_(True)
    with store.as_default():
        out = PointNonLocalCell(f, p, [32, 24], True, 0.9, None, "nlc", True, mode="concat")
    assert np.abs(out.detach().double().numpy() - g["out"]).max() <= 1e-5 * np.abs(g["out"]).max()
    (out * torch.from_numpy(probe_weights("nlc/out", out.shape)).float()).sum().backward()
    scale = max(float(g[k][1]) for k in g.files if k.startswith("grad/"))
    named = {n: v for n, v in store.variables().items() if v.requires_grad}
    named.update({"in/feature": f, "in/new_point": p})
    assert {k[5:] for k in g.files if k.startswith("grad/")} == set(named)
    for n, v in named.items():
        want = g["grad/" + n]
        got = digest(n, v.grad.double().numpy())
        assert np.abs(got - want).max() <= (1e-4 * want[1] + 1e-6 * scale) * max(1.0, np.sqrt(v.numel())), n
    for k in g.files:
        if k.startswith("moving/"):
            np.testing.assert_allclose(store.variables()[k[7:]].numpy(), g[k], rtol=1e-4, atol=1e-6)


def test_conv2d_pair_creates_the_variables_of_two_conv2d_calls_and_matches_them_on_cpu():
    """tf_util.conv2d_pair (the fused conv0 -> conv1 pair of a local cell on CUDA) must be indistinguishable from two
    conv2d calls: same variable names in the same creation order, same values, same moving averages."""
    from pointasnl_b200 import tf_util
    x = torch.randn(2, 5, 4, 7)
    outs, names, moving = [], [], []
    for pair in (False, True):
        store = tf_util.VariableStore("cpu", init_fn=lambda full, shape, leaf: golden_init(full, shape, leaf))
        with store.as_default(), tf_util.variable_scope("layer"):
            if pair:
                y = tf_util.conv2d_pair(x, [6, 5], ("conv0", "conv1"), bn_decay=0.9, is_training=True, weight_decay=0.1)
            else:
                y = x
                for i, c in enumerate([6, 5]):
                    y = tf_util.conv2d(y, c, [1, 1], padding="VALID", stride=[1, 1], bn=True, is_training=True,
                                       scope="conv%d" % i, bn_decay=0.9, weight_decay=0.1)
        outs.append(y)
        names.append(list(store.variables()))
        moving.append({n: v.clone() for n, v in store.variables().items() if "moving" in n})
    assert names[0] == names[1]
    assert names[0][:2] == ["layer/conv0/weights", "layer/conv0/biases"] and "layer/conv1/bn/moving_variance" in names[0]
    torch.testing.assert_close(outs[0], outs[1], rtol=0, atol=0)
    for n in moving[0]:
        torch.testing.assert_close(moving[0][n], moving[1][n], rtol=0, atol=0)
