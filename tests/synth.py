"""Seeded synthetic inputs.  numpy.random.default_rng(seed), float32 everywhere."""
import numpy as np


def pc_normalize(pc):
    """centroid-subtract, scale to the unit sphere (reference modelnet_dataset.py:9-14)."""
    pc = pc - pc.mean(axis=0, keepdims=True)
    m = np.sqrt((pc ** 2).sum(axis=1)).max()
    return (pc / m).astype(np.float32)


def uniform_clouds(b, n, seed=0, normalize=True):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, (b, n, 3)).astype(np.float32)
    if normalize:
        x = np.stack([pc_normalize(c) for c in x])
    return x


def lattice_clouds(b, n, side=4, seed=0):
    """integer-lattice coordinates: masses of exact distance ties."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, side, (b, n, 3)).astype(np.float32)


def duplicated_clouds(b, n, frac=0.05, seed=0):
    """uniform cloud in which a fraction of points are exact copies of other points (the ScanNet/KITTI pipelines pad
    clouds by duplicating points, reference ScanNet/scannet_dataset_grid.py:692-703) and point 0 is copied into
    dropped slots (provider.py:246-254)."""
    rng = np.random.default_rng(seed)
    x = uniform_clouds(b, n, seed, normalize=True)
    k = max(1, int(n * frac))
    for c in x:
        dst = rng.choice(n, k, replace=False)
        src = rng.choice(n, k, replace=True)
        c[dst] = c[src]
        c[rng.choice(n, max(1, k // 4), replace=False)] = c[0]
    return x


def modelnet_batch(b=32, n=1024, seed=1):
    """config 2: xyz on random-pose unit sphere / box surfaces -> pc_normalize -> train-time jitter (scale U[0.8,1.25],
    shift U[-0.1,0.1], reference provider.py:219-244), normals = normalised Gaussian.  (b, n, 6), labels (b,)."""
    rng = np.random.default_rng(seed)
    out = np.empty((b, n, 6), np.float32)
    for i in range(b):
        if i % 2 == 0:
            p = rng.normal(size=(n, 3))
            p /= np.linalg.norm(p, axis=1, keepdims=True)
            p *= rng.uniform(0.5, 1.0, (1, 3))
        else:
            p = rng.uniform(-1, 1, (n, 3))
            ax = rng.integers(0, 3, n)
            p[np.arange(n), ax] = np.sign(p[np.arange(n), ax])
            p *= rng.uniform(0.4, 1.0, (1, 3))
        a = rng.uniform(0, 2 * np.pi)
        rot = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        p = pc_normalize(p @ rot.T)
        p = p * rng.uniform(0.8, 1.25) + rng.uniform(-0.1, 0.1, (1, 3))
        nrm = rng.normal(size=(n, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        out[i, :, :3] = p
        out[i, :, 3:] = nrm
    labels = rng.integers(0, 40, b).astype(np.int64)
    return out, labels


def scannet_batch(b=4, n=10240, seed=2, dup_frac=0.0):
    """config 4: 0.04 m voxel-quantised-then-jittered points in a sphere of radius 2 m, xyz + rgb in [0,1]."""
    rng = np.random.default_rng(seed)
    pts = np.empty((b, n, 6), np.float32)
    for i in range(b):
        p = np.empty((0, 3))
        while len(p) < n:
            c = rng.uniform(-2, 2, (4 * n, 3))
            c = c[(c ** 2).sum(1) < 4.0]
            c = np.unique(np.floor(c / 0.04), axis=0) * 0.04
            p = np.concatenate([p, c])
        p = p[rng.permutation(len(p))[:n]]
        p = p * rng.uniform(0.9, 1.1, (1, 3)) + rng.normal(0, 0.001, (n, 3))
        if dup_frac > 0:
            k = int(n * dup_frac)
            p[rng.choice(n, k, replace=False)] = p[rng.choice(n, k, replace=True)]
        pts[i, :, :3] = p
        pts[i, :, 3:] = rng.uniform(0, 1, (n, 3))
    labels = rng.integers(0, 21, (b, n)).astype(np.int64)
    smpw = rng.uniform(1.0, 4.9, (b, n)).astype(np.float32)
    return pts, labels, smpw


def kitti_batch(b=2, n=65536, seed=3):
    """config 5: LiDAR-like rings, range U[2,50] m, 64 elevation rings in [-25,+3] degrees, xyz + remission; not centred."""
    rng = np.random.default_rng(seed)
    pts = np.empty((b, n, 4), np.float32)
    for i in range(b):
        r = rng.uniform(2, 50, n)
        az = rng.uniform(0, 2 * np.pi, n)
        el = np.deg2rad(-25 + 28 * rng.integers(0, 64, n) / 63.0)
        p = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], 1)
        pts[i, :, :3] = np.floor(p / 0.06) * 0.06 + rng.uniform(0, 0.06, (n, 3))
        pts[i, :, 3] = rng.uniform(0, 1, n)
    labels = rng.integers(0, 20, (b, n)).astype(np.int64)
    return pts, labels
