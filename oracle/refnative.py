"""Bindings of the REFERENCE's own code, compiled untouched into oracle/_ref/ by oracle/build_ref.sh.

Test infrastructure (see oracle/__init__.py).  Nothing here reads /root/reference at run time: only the
prebuilt .so files are loaded, so this works on the GPU box too.

    CPU:  knn_batch_omp (knn_.cxx:104-135), threenn_cpu / threeinterpolate_cpu / threeinterpolate_grad_cpu
          (tf_interpolate.cpp:60-153)                                       -> numpy in / numpy out
    GPU:  farthestpointsamplingLauncher, gatherpointLauncher, scatteraddpointLauncher (tf_sampling_g.cu:194-211),
          queryBallPointLauncher, groupPointLauncher, groupPointGradLauncher, selectionSortLauncher
          (tf_grouping_g.cu:125-141)                                        -> torch CUDA tensors in / out.
          These launch on the legacy default stream, exactly as the reference does; the wrappers
          synchronise the device before and after.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_libs = {}

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_l = ctypes.POINTER(ctypes.c_long)


def available(name):
    return os.path.exists(os.path.join(REF_DIR, name))


def _lib(name):
    if name not in _libs:
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run oracle/build_ref.sh where /root/reference is mounted")
        _libs[name] = ctypes.CDLL(path)
    return _libs[name]


# ---------------------------------------------------------------- CPU pieces
def knn_batch(pts, queries, k, threads=None):
    """nearest_neighbors.knn_batch(pts, queries, K, omp=True) (knn.pyx:71-109) minus the Cython shim."""
    pts = np.ascontiguousarray(pts, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    b, n, dim = pts.shape
    m = queries.shape[1]
    out = np.zeros((b, m, k), np.int64)
    lib = _lib("ref_knn.so")
    fn = lib._Z17cpp_knn_batch_ompPKfmmmS0_mmPl
    fn.restype = None
    fn.argtypes = [_f, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _f, ctypes.c_size_t, ctypes.c_size_t, _l]
    if threads is not None:
        # libgomp is shared by every OpenMP user in the process
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    fn(pts.ctypes.data_as(_f), b, n, dim, queries.ctypes.data_as(_f), m, k, out.ctypes.data_as(_l))
    return out


def three_nn(xyz1, xyz2):
    xyz1 = np.ascontiguousarray(xyz1, np.float32)
    xyz2 = np.ascontiguousarray(xyz2, np.float32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    _lib("ref_interp.so").threenn_cpu(b, n, m, xyz1.ctypes.data_as(_f), xyz2.ctypes.data_as(_f),
                                      dist.ctypes.data_as(_f), idx.ctypes.data_as(_i))
    return dist, idx


def three_interpolate(points, idx, weight):
    points = np.ascontiguousarray(points, np.float32)
    idx = np.ascontiguousarray(idx, np.int32)
    weight = np.ascontiguousarray(weight, np.float32)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    _lib("ref_interp.so").threeinterpolate_cpu(b, m, c, n, points.ctypes.data_as(_f), idx.ctypes.data_as(_i),
                                               weight.ctypes.data_as(_f), out.ctypes.data_as(_f))
    return out


def three_interpolate_grad(m, idx, weight, grad_out):
    idx = np.ascontiguousarray(idx, np.int32)
    weight = np.ascontiguousarray(weight, np.float32)
    grad_out = np.ascontiguousarray(grad_out, np.float32)
    b, n, c = grad_out.shape
    g = np.zeros((b, m, c), np.float32)  # the TF op memsets before calling (tf_interpolate.cpp:258)
    _lib("ref_interp.so").threeinterpolate_grad_cpu(b, n, c, m, grad_out.ctypes.data_as(_f), idx.ctypes.data_as(_i),
                                                    weight.ctypes.data_as(_f), g.ctypes.data_as(_f))
    return g


def grid_subsampling(points, features=None, classes=None, dl=0.1):
    """The reference's grid_subsampling() (grid_subsampling.cpp:5-106) through oracle/ref_shims; outputs in the
    reference's own unordered_map iteration order."""
    pts = np.ascontiguousarray(points, np.float32)
    n = pts.shape[0]
    feats = np.ascontiguousarray(features, np.float32) if features is not None else None
    cls = np.ascontiguousarray(classes, np.int32).reshape(n, -1) if classes is not None else None
    fdim = 0 if feats is None else feats.shape[1]
    ldim = 0 if cls is None else cls.shape[1]
    op = np.empty((n, 3), np.float32)
    of = np.empty((n, max(fdim, 1)), np.float32)
    oc = np.empty((n, max(ldim, 1)), np.int32)
    fn = _lib("ref_gridsub.so").ref_grid_subsampling
    fn.restype = ctypes.c_int
    m = fn(n, pts.ctypes.data_as(_f), fdim, feats.ctypes.data_as(_f) if feats is not None else None, ldim,
           cls.ctypes.data_as(_i) if cls is not None else None, ctypes.c_float(dl), op.ctypes.data_as(_f),
           of.ctypes.data_as(_f), oc.ctypes.data_as(_i))
    out = [op[:m]]
    if fdim:
        out.append(of[:m].reshape(m, fdim))
    if ldim:
        out.append(oc[:m].reshape(m, ldim))
    return out[0] if len(out) == 1 else tuple(out)


# ---------------------------------------------------------------- GPU pieces (torch CUDA tensors)
def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _sync():
    import torch
    torch.cuda.synchronize()


def gpu_fps(xyz, m):
    import torch
    b, n, _ = xyz.shape
    xyz = xyz.contiguous().float()
    temp = torch.empty((32, n), device=xyz.device, dtype=torch.float32)  # tf_sampling.cpp:115
    out = torch.zeros((b, m), device=xyz.device, dtype=torch.int32)
    _sync()
    _lib("ref_sampling.so")._Z29farthestpointsamplingLauncheriiiPKfPfPi(b, n, m, _p(xyz), _p(temp), _p(out))
    _sync()
    return out


def gpu_gather_point(inp, idx):
    import torch
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp = inp.contiguous().float()
    idx = idx.contiguous().int()
    out = torch.empty((b, m, 3), device=inp.device, dtype=torch.float32)
    _sync()
    _lib("ref_sampling.so")._Z19gatherpointLauncheriiiPKfPKiPf(b, n, m, _p(inp), _p(idx), _p(out))
    _sync()
    return out


def gpu_gather_point_grad(n, idx, out_g):
    import torch
    b, m = idx.shape
    idx = idx.contiguous().int()
    out_g = out_g.contiguous().float()
    g = torch.zeros((b, n, 3), device=idx.device, dtype=torch.float32)  # tf_sampling.cpp:174
    _sync()
    _lib("ref_sampling.so")._Z23scatteraddpointLauncheriiiPKfPKiPf(b, n, m, _p(out_g), _p(idx), _p(g))
    _sync()
    return g


def gpu_ball_query(radius, nsample, xyz1, xyz2):
    import torch
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    # zero-initialised so that no-hit rows are comparable (the reference leaves them uninitialised)
    idx = torch.zeros((b, m, nsample), device=xyz1.device, dtype=torch.int32)
    cnt = torch.zeros((b, m), device=xyz1.device, dtype=torch.int32)
    _sync()
    fn = _lib("ref_grouping.so")._Z22queryBallPointLauncheriiifiPKfS0_PiS1_
    fn(b, n, m, ctypes.c_float(radius), nsample, _p(xyz1), _p(xyz2), _p(idx), _p(cnt))
    _sync()
    return idx, cnt


def gpu_group_point(points, idx):
    import torch
    b, n, c = points.shape
    _, m, ns = idx.shape
    points = points.contiguous().float()
    idx = idx.contiguous().int()
    out = torch.empty((b, m, ns, c), device=points.device, dtype=torch.float32)
    _sync()
    _lib("ref_grouping.so")._Z18groupPointLauncheriiiiiPKfPKiPf(b, n, c, m, ns, _p(points), _p(idx), _p(out))
    _sync()
    return out


def gpu_group_point_grad(n, idx, grad_out):
    import torch
    b, m, ns, c = grad_out.shape
    idx = idx.contiguous().int()
    grad_out = grad_out.contiguous().float()
    g = torch.zeros((b, n, c), device=idx.device, dtype=torch.float32)  # tf_grouping.cpp:204
    _sync()
    _lib("ref_grouping.so")._Z22groupPointGradLauncheriiiiiPKfPKiPf(b, n, c, m, ns, _p(grad_out), _p(idx), _p(g))
    _sync()
    return g


def gpu_selection_sort(k, dist):
    import torch
    b, m, n = dist.shape
    dist = dist.contiguous().float()
    outi = torch.empty((b, m, n), device=dist.device, dtype=torch.int32)
    out = torch.empty((b, m, n), device=dist.device, dtype=torch.float32)
    _sync()
    _lib("ref_grouping.so")._Z21selectionSortLauncheriiiiPKfPiPf(b, n, m, k, _p(dist), _p(outi), _p(out))
    _sync()
    return outi, out


def gpu_prob_sample(inp, inpr):
    """probsampleLauncher (tf_sampling_g.cu:194-197) -> (indices, running sums), torch CUDA tensors."""
    import torch
    b, n = inp.shape
    m = inpr.shape[1]
    inp = inp.contiguous().float()
    inpr = inpr.contiguous().float()
    temp = torch.empty((b, n), device=inp.device, dtype=torch.float32)
    out = torch.zeros((b, m), device=inp.device, dtype=torch.int32)
    _sync()
    _lib("ref_sampling.so")._Z18probsampleLauncheriiiPKfS0_PfPi(b, n, m, _p(inp), _p(inpr), _p(temp), _p(out))
    _sync()
    return out, temp
