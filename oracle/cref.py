"""numpy bindings of oracle/cpu_ref.c (test infrastructure; see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_max_threads.restype = ctypes.c_int
        _LIB.oracle_get_threads.restype = ctypes.c_int
    return _LIB


def set_threads(t):
    lib().oracle_set_threads(int(t))


def max_threads():
    return int(lib().oracle_max_threads())


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fps(xyz, m, sim512=False):
    xyz = _f32(xyz)
    b, n, _ = xyz.shape
    out = np.zeros((b, m), np.int32)
    fn = lib().oracle_fps_sim512 if sim512 else lib().oracle_fps
    fn(b, n, m, _fp(xyz), _ip(out))
    return out


def gather_point(inp, idx):
    inp, idx = _f32(inp), _i32(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.empty((b, m, 3), np.float32)
    lib().oracle_gather_point(b, n, m, _fp(inp), _ip(idx), _fp(out))
    return out


def gather_point_grad(n, idx, out_g):
    idx, out_g = _i32(idx), _f32(out_g)
    b, m = idx.shape
    g = np.empty((b, n, 3), np.float32)
    lib().oracle_gather_point_grad(b, n, m, _fp(out_g), _ip(idx), _fp(g))
    return g


def ball_query(radius, nsample, xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.empty((b, m, nsample), np.int32)
    cnt = np.empty((b, m), np.int32)
    lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), nsample, _fp(xyz1), _fp(xyz2), _ip(idx), _ip(cnt))
    return idx, cnt


def group_point(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    lib().oracle_group_point(b, n, c, m, ns, _fp(points), _ip(idx), _fp(out))
    return out


def group_point_grad(n, idx, grad_out):
    idx, grad_out = _i32(idx), _f32(grad_out)
    b, m, ns, c = grad_out.shape
    g = np.empty((b, n, c), np.float32)
    lib().oracle_group_point_grad(b, n, c, m, ns, _fp(grad_out), _ip(idx), _fp(g))
    return g


def knn(k, xyz1, xyz2, return_dist=False):
    """k nearest support points (xyz1) of every query (xyz2); ascending (distance, index)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.empty((b, m, k), np.int32)
    dist = np.empty((b, m, k), np.float32)
    lib().oracle_knn(b, n, m, k, _fp(xyz1), _fp(xyz2), _ip(idx), _fp(dist))
    return (dist, idx) if return_dist else idx


def selection_sort(k, dist):
    dist = _f32(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    lib().oracle_selection_sort(b, n, m, k, _fp(dist), _ip(outi), _fp(out))
    return outi, out


def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().oracle_three_nn(b, n, m, _fp(xyz1), _fp(xyz2), _fp(dist), _ip(idx))
    return dist, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    lib().oracle_three_interpolate(b, m, c, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(m, idx, weight, grad_out):
    idx, weight, grad_out = _i32(idx), _f32(weight), _f32(grad_out)
    b, n, c = grad_out.shape
    g = np.empty((b, m, c), np.float32)
    lib().oracle_three_interpolate_grad(b, n, c, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def grid_subsample(points, features=None, classes=None, dl=0.1, tie_largest=False):
    """oracle_grid_subsample (grid_subsampling.cpp:5-106): -> sub_points [, sub_features] [, sub_classes]; cells in order
    of first appearance, label ties -> smallest label."""
    pts = _f32(points)
    n = pts.shape[0]
    feats = _f32(features) if features is not None else None
    cls = _i32(classes).reshape(n, -1) if classes is not None else None
    fdim = 0 if feats is None else feats.shape[1]
    ldim = 0 if cls is None else cls.shape[1]
    op = np.empty((n, 3), np.float32)
    of = np.empty((n, max(fdim, 1)), np.float32)
    oc = np.empty((n, max(ldim, 1)), np.int32)
    fn = lib().oracle_grid_subsample_ex
    fn.restype = ctypes.c_int
    m = fn(n, _fp(pts), fdim, _fp(feats) if feats is not None else None, ldim, _ip(cls) if cls is not None else None,
           ctypes.c_float(dl), _fp(op), _fp(of), _ip(oc), 1 if tie_largest else 0)
    out = [op[:m]]
    if fdim:
        out.append(of[:m, :fdim])
    if ldim:
        out.append(oc[:m, :ldim])
    return out[0] if len(out) == 1 else tuple(out)


def prob_sample(inp, inpr):
    """oracle_prob_sample (tf_sampling_g.cu:7-104) -> (indices (b,m) int32, running sums (b,n) float32)."""
    inp, inpr = _f32(inp), _f32(inpr)
    b, n = inp.shape
    m = inpr.shape[1]
    temp = np.empty((b, n), np.float32)
    out = np.empty((b, m), np.int32)
    lib().oracle_prob_sample(b, n, m, _fp(inp), _fp(inpr), _fp(temp), _ip(out))
    return out, temp
