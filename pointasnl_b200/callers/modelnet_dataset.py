"""ModelNet40 (normal-resampled) loader with the reference's interface (modelnet_dataset.py:38-137): the same directory
layout (`modelnet40_shape_names.txt`, `modelnet40_train.txt` / `_test.txt`, `<shape>/<shape>_NNNN.txt` with
x,y,z,nx,ny,nz per line), `reset()` / `has_next_batch()` / `next_batch()` and per-cloud `pc_normalize`
(centroid-subtract, scale to the unit sphere, :9-14).  Host-side numpy; the batches feed TrainStep.set_inputs."""
import os

import numpy as np


def pc_normalize(pc):
    pc = pc - pc.mean(axis=0)
    return pc / np.sqrt((pc ** 2).sum(axis=1)).max()


class ModelNetDataset:
    def __init__(self, root, batch_size=32, npoints=1024, split="train", normalize=True, normal_channel=False,
                 modelnet10=False, cache_size=15000, shuffle=None, seed=None):
        if split not in ("train", "test"):
            raise ValueError("split must be 'train' or 'test'")
        tag = "modelnet10" if modelnet10 else "modelnet40"
        self.root, self.batch_size, self.npoints = root, batch_size, npoints
        self.normalize, self.normal_channel = normalize, normal_channel
        names = [ln.rstrip() for ln in open(os.path.join(root, f"{tag}_shape_names.txt"))]
        self.classes = {n: i for i, n in enumerate(names)}
        ids = [ln.rstrip() for ln in open(os.path.join(root, f"{tag}_{split}.txt"))]
        self.datapath = [("_".join(i.split("_")[:-1]), os.path.join(root, "_".join(i.split("_")[:-1]), i + ".txt")) for i in ids]
        self.cache, self.cache_size = {}, cache_size
        self.shuffle = (split == "train") if shuffle is None else shuffle
        self.rng = np.random.default_rng(seed)
        self.reset()

    def __len__(self):
        return len(self.datapath)

    def num_channel(self):
        return 6 if self.normal_channel else 3

    def __getitem__(self, index):
        if index not in self.cache:
            shape, path = self.datapath[index]
            pts = np.loadtxt(path, delimiter=",").astype(np.float32)[: self.npoints]
            if self.normalize:
                pts[:, 0:3] = pc_normalize(pts[:, 0:3])
            if not self.normal_channel:
                pts = pts[:, 0:3]
            item = (pts, np.int32(self.classes[shape]))
            if len(self.cache) >= self.cache_size:
                return item
            self.cache[index] = item
        return self.cache[index]

    def reset(self):
        self.idxs = np.arange(len(self.datapath))
        if self.shuffle:
            self.rng.shuffle(self.idxs)
        self.num_batches = (len(self.datapath) + self.batch_size - 1) // self.batch_size
        self.batch_idx = 0

    def has_next_batch(self):
        return self.batch_idx < self.num_batches

    def next_batch(self):
        """(bsize, npoints, 3|6) float32, (bsize,) int32; the last batch may be smaller than batch_size."""
        lo = self.batch_idx * self.batch_size
        hi = min(lo + self.batch_size, len(self.datapath))
        data = np.zeros((hi - lo, self.npoints, self.num_channel()), np.float32)
        label = np.zeros((hi - lo,), np.int32)
        for i in range(lo, hi):
            data[i - lo], label[i - lo] = self[self.idxs[i]]
        self.batch_idx += 1
        return data, label
