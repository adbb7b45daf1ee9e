"""Host-side logic on the CPU: the TF-style variable store, op argument validation, the data-parallel gradient
bucket (world_size-2 gloo), the oracle model restatement."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointasnl_b200 import dp, tf_util


def test_variable_store_names_follow_tf_scopes_and_reuse():
    store = tf_util.VariableStore("cpu", seed=0)
    x = torch.randn(2, 5, 7, 6)
    with store.as_default():
        with tf_util.variable_scope("layer1"):
            with tf_util.variable_scope("layer1"):
                y = tf_util.conv2d(x, 8, [1, 1], "conv_kv", padding="VALID", bn=True, is_training=True,
                                   activation_fn=None)
                y2 = tf_util.conv2d(x, 8, [1, 1], "conv_kv", padding="VALID", bn=True, is_training=True,
                                    activation_fn=None)
    names = set(store.variables())
    assert names == {"layer1/layer1/conv_kv/weights", "layer1/layer1/conv_kv/biases", "layer1/layer1/conv_kv/bn/beta",
                     "layer1/layer1/conv_kv/bn/gamma", "layer1/layer1/conv_kv/bn/moving_mean",
                     "layer1/layer1/conv_kv/bn/moving_variance"}
    assert torch.equal(y, y2)                                   # AUTO_REUSE: the second call reuses the variables
    assert tuple(store.variables()["layer1/layer1/conv_kv/weights"].shape) == (1, 1, 6, 8)
    assert len(list(store.parameters())) == 4                   # moving stats are buffers, not parameters
    with pytest.raises(ValueError, match="exists with shape"):
        with store.as_default(), tf_util.variable_scope("layer1"), tf_util.variable_scope("layer1"):
            tf_util.conv2d(torch.randn(2, 5, 7, 9), 8, [1, 1], "conv_kv", padding="VALID")


def test_layers_follow_bias_bn_activation_order_and_tf_bn_semantics():
    store = tf_util.VariableStore("cpu", seed=1)
    x = torch.randn(3, 11, 5)
    with store.as_default():
        y = tf_util.conv1d(x, 4, 1, "c", padding="VALID", bn=True, is_training=True)
    v = store.variables()
    z = x @ v["c/weights"][0] + v["c/biases"]
    mean, var = z.mean((0, 1)), z.var((0, 1), unbiased=False)
    want = torch.relu((z - mean) / torch.sqrt(var + 1e-3) * v["c/bn/gamma"] + v["c/bn/beta"])   # eps 1e-3, biased var
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-6)
    # full-width VALID conv == matmul over the flattened (W, C) axes
    with store.as_default():
        t = torch.randn(2, 3, 6, 5)
        out = tf_util.conv2d(t, 7, [1, 6], "after", padding="VALID", activation_fn=None)
    w = store.variables()["after/weights"]
    ref = torch.einsum("bhwc,wco->bho", t, w[0]) + store.variables()["after/biases"]
    torch.testing.assert_close(out.squeeze(2), ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError, match="no VariableStore"):
        tf_util.fully_connected(torch.randn(2, 3), 4, "fc")


def test_ops_refuse_cpu_tensors_there_is_no_fallback():
    from pointasnl_b200 import pointasnl_util, tf_grouping, tf_interpolate, tf_sampling
    x = torch.zeros(1, 16, 3)
    for fn in (lambda: tf_sampling.farthest_point_sample(4, x), lambda: tf_grouping.knn_point(2, x, x),
               lambda: tf_grouping.query_ball_point(0.1, 2, x, x), lambda: tf_interpolate.three_nn(x, x),
               lambda: pointasnl_util.knn_query(2, x, x), lambda: pointasnl_util.sampling(4, x),
               lambda: tf_sampling.prob_sample(x, x)):
        with pytest.raises(ValueError, match="no CPU path"):
            fn()


def test_shard_range_partitions_clouds():
    for total, world in [(32, 8), (32, 3), (4, 8), (7, 2)]:
        spans = [dp.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    bucket = dp.FlatGradBucket(model.parameters())
    gen = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 6, generator=gen), torch.randn(8, 3, generator=gen)
    lo, hi = dp.shard_range(8, rank, world)
    bucket.zero()
    ((model(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean().backward()
    bucket.check_aliasing()
    accumulated = bucket.flat.clone()
    # the path TrainStep takes: autograd assigns fresh gradients, one multi-tensor copy packs them into the bucket
    bucket.detach()
    ((model(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean().backward()
    bucket.pack()
    bucket.check_aliasing()
    assert torch.equal(bucket.flat, accumulated)
    bucket.all_reduce_mean()
    slowest = dp.max_over_ranks(float(rank + 1), torch.device("cpu"))
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        # mean over ranks of per-rank means == mean over the concatenated batch (equal shard sizes)
        ((ref(x_all) - y_all) ** 2).mean().backward()
        want = torch.cat([p.grad.flatten() for p in ref.parameters()])
        out.put((torch.allclose(bucket.flat, want, rtol=1e-5, atol=1e-6), slowest, bucket.nbytes()))
    dist.destroy_process_group()


def test_dp_bucket_single_allreduce_matches_full_batch_gradients_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, slowest, nbytes = out.get()
    assert ok and slowest == 2.0 and nbytes == (6 * 5 + 5 + 5 * 3 + 3) * 4


def test_oracle_model_restatement_runs_and_has_the_reference_parameter_count():
    from oracle import model_ref
    from tests import synth
    pc, labels = synth.modelnet_batch(2, 1024, 1)
    r = model_ref.Ref(None, dropout=False)
    net, total = model_ref.cls_forward(r, torch.from_numpy(pc), torch.from_numpy(labels))
    total.backward()
    assert tuple(net.shape) == (2, 40) and torch.isfinite(total)
    n_params = sum(v.numel() for v in r.v.values())
    assert n_params == 3359368            # "3.36 M without AS" (weights + biases + BN gamma/beta)
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in r.v.values())
    assert "layer1/layer1/conv_kv/weights" in r.v and "layer2/after_conv/weights" in r.v
    r2 = model_ref.Ref(None, dropout=False)
    model_ref.cls_forward(r2, torch.from_numpy(pc), torch.from_numpy(labels), adaptive_sample=True)
    assert "layer1/layer1/layer1/conv_kv_ds/weights" in r2.v      # the triple-nested scope of the reference


def test_zero_arena_hands_out_disjoint_zeroed_slices_and_grows_on_reset():
    """Host logic of the per-step zero arena used by the dense glue (pointasnl_b200/_ops.py: ZeroArena)."""
    from pointasnl_b200 import _ops
    arena = _ops.ZeroArena("cpu")
    arena.reset()
    assert arena.zeros((4, 3), torch.float32) is None            # nothing reserved yet: callers fall back to torch.zeros
    assert arena.zeros((5,), torch.float64) is None
    arena.reset()                                                # grows to the high-water mark of the previous step
    a = arena.zeros((4, 3), torch.float32)
    b = arena.zeros((5,), torch.float64)
    assert a is not None and b is not None and a.shape == (4, 3) and b.dtype == torch.float64
    assert float(a.abs().sum()) == 0.0 and float(b.abs().sum()) == 0.0
    a.fill_(1.0)
    b.fill_(2.0)
    assert float(a.sum()) == 12.0 and float(b.sum()) == 10.0     # disjoint storage
    assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
    assert arena.zeros((1000,), torch.float32) is None           # over capacity: fall back, remembered for next reset
    arena.reset()
    c = arena.zeros((4, 3), torch.float32)
    assert float(c.abs().sum()) == 0.0                           # cleared again
    with _ops.use_zero_arena(arena):
        t, from_arena = _ops._zeros((5,), torch.float64, "cpu")
        assert from_arena and float(t.abs().sum()) == 0.0
    t, from_arena = _ops._zeros((5,), torch.float64, "cpu")
    assert not from_arena


def test_flat_parameter_adam_matches_per_tensor_adam():
    """FlatGradBucket.flatten_parameters: one Adam over the flat buffer == Adam over the individual tensors."""
    torch.manual_seed(3)
    m1 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    m2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    m2.load_state_dict(m1.state_dict())
    x, y = torch.randn(16, 6), torch.randn(16, 3)
    bucket = dp.FlatGradBucket(m1.parameters())
    fp = bucket.flatten_parameters()
    o1 = torch.optim.Adam([fp], lr=1e-2)
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-2)
    for _ in range(3):
        bucket.detach()
        ((m1(x) - y) ** 2).mean().backward()
        bucket.pack()
        assert fp.grad.data_ptr() == bucket.flat.data_ptr()
        o1.step()
        o2.zero_grad()
        ((m2(x) - y) ** 2).mean().backward()
        o2.step()
    for a, b in zip(m1.parameters(), m2.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
