"""TF V2 checkpoint reader / writer (pointasnl_b200/tf_checkpoint.py) and the ModelNet loader mirror: no GPU.
The reader cannot be checked against a file TensorFlow wrote (TF is not installable here): known-answer tests of the
pieces the format fixes (crc32c, varints, block prefix compression, footer) + round trips through a VariableStore."""
import os
import struct

import numpy as np
import pytest
import torch

from pointasnl_b200 import tf_checkpoint as ck
from pointasnl_b200 import tf_util


def test_crc32c_known_answers_and_mask():
    assert ck.crc32c(b"123456789") == 0xE3069283                       # the standard CRC-32C check value
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA                        # RFC 3720 B.4 test vector
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck._mask_crc(0) == 0xA282EAD8


def test_block_prefix_compression_round_trip():
    pairs = [(f"layer1/conv{i}/weights".encode(), bytes([i]) * (i % 5 + 1)) for i in range(40)]
    pairs.sort()
    assert ck._block_entries(ck._build_block(pairs, restart_interval=4)) == pairs
    blk = ck._build_block(pairs, restart_interval=16)
    assert len(blk) < sum(len(k) + len(v) + 3 for k, v in pairs)        # shared prefixes were elided


def test_checkpoint_round_trip_many_blocks(tmp_path):
    rng = np.random.default_rng(0)
    variables = {f"layer{i // 7}/conv{i % 7}/weights": rng.normal(size=(1, 1, i % 9 + 1, 4)).astype(np.float32)
                 for i in range(300)}
    variables["global_step"] = np.array(7, np.int64)
    variables["fc1/weights/Adam"] = np.zeros((3, 3), np.float32)
    prefix = str(tmp_path / "model.ckpt")
    ck.save_checkpoint(prefix, variables, entries_per_block=17)
    header, entries = ck.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and set(entries) == set(variables)
    got = ck.load_checkpoint(prefix)
    assert "global_step" not in got and "fc1/weights/Adam" not in got   # optimizer state is not model state
    for n, v in variables.items():
        if n in got:
            assert got[n].dtype == v.dtype and np.array_equal(got[n], v)
    raw = open(prefix + ".data-00000-of-00001", "rb").read()
    e = entries["layer0/conv0/weights"]
    assert ck._mask_crc(ck.crc32c(raw[e["offset"]:e["offset"] + e["size"]])) == e["crc32c"]
    assert struct.unpack("<Q", open(prefix + ".index", "rb").read()[-8:])[0] == ck.TABLE_MAGIC


def test_store_save_restore_by_reference_variable_names(tmp_path):
    def build(store):
        x = torch.randn(2, 5, 4, 6)
        with store.as_default():
            y = tf_util.conv2d(x, 8, [1, 1], "layer1/conv0", padding="VALID", bn=True, is_training=True, bn_decay=0.9)
            y = tf_util.conv1d(y.reshape(2, 20, 8), 5, 1, "layer1/skip", padding="VALID", bn=True, is_training=True)
            tf_util.fully_connected(y.reshape(40, 5), 3, "fc3", activation_fn=None)
    a = tf_util.VariableStore("cpu", seed=1)
    build(a)
    prefix = str(tmp_path / "best_model.ckpt")
    ck.save(a, prefix)
    names = set(ck.load_checkpoint(prefix))
    assert {"layer1/conv0/weights", "layer1/conv0/biases", "layer1/conv0/bn/gamma", "layer1/conv0/bn/moving_variance",
            "layer1/skip/weights", "fc3/weights", "fc3/biases"} <= names
    b = tf_util.VariableStore("cpu", seed=2, init_fn=ck.initializer_from_checkpoint(prefix))
    build(b)                                                               # variables take the checkpoint values at creation
    c = tf_util.VariableStore("cpu", seed=3)
    build(c)
    assert ck.restore(c, prefix) == []
    for n, v in a.variables().items():
        if "moving_" in n:
            continue                                                       # build() ran one more training-mode update on b
        assert torch.equal(b.variables()[n], v) and torch.equal(c.variables()[n], v)
    assert torch.equal(c.variables()["layer1/conv0/bn/moving_mean"], a.variables()["layer1/conv0/bn/moving_mean"])
    with pytest.raises(ValueError):
        ck.read_index(prefix + ".data-00000-of-00001")                     # not a table: bad magic


def test_modelnet_loader_interface(tmp_path):
    from pointasnl_b200.callers.modelnet_dataset import ModelNetDataset, pc_normalize
    root = tmp_path / "modelnet40_normal_resampled"
    rng = np.random.default_rng(0)
    names = ["chair", "night_stand"]
    (root).mkdir()
    open(root / "modelnet40_shape_names.txt", "w").write("\n".join(names) + "\n")
    ids = {"train": ["chair_0001", "chair_0002", "night_stand_0001"], "test": ["night_stand_0002"]}
    for split, lst in ids.items():
        open(root / f"modelnet40_{split}.txt", "w").write("\n".join(lst) + "\n")
        for i in lst:
            d = root / "_".join(i.split("_")[:-1])
            d.mkdir(exist_ok=True)
            np.savetxt(d / (i + ".txt"), rng.normal(size=(50, 6)), delimiter=",", fmt="%.6f")
    ds = ModelNetDataset(str(root), batch_size=2, npoints=32, split="train", normal_channel=True, seed=0)
    assert len(ds) == 3 and ds.num_channel() == 6 and ds.shuffle
    seen = 0
    while ds.has_next_batch():
        data, label = ds.next_batch()
        assert data.shape[1:] == (32, 6) and data.dtype == np.float32 and label.dtype == np.int32
        assert np.allclose(np.sqrt((data[..., :3] ** 2).sum(-1)).max(axis=1), 1.0, atol=1e-5)     # unit sphere
        seen += len(label)
    assert seen == 3
    test = ModelNetDataset(str(root), batch_size=2, npoints=32, split="test")
    d, lab = test.next_batch()
    assert not test.shuffle and d.shape == (1, 32, 3) and int(lab[0]) == 1
    p = rng.normal(size=(10, 3))
    assert abs(np.sqrt((pc_normalize(p) ** 2).sum(1)).max() - 1) < 1e-12
