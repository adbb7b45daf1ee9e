"""Drop-in for the reference's grid subsampling entry point (ScanNet/scannet_dataset_grid.py:20-38 ->
cpp_wrappers.cpp_subsampling.grid_subsampling.compute -> grid_subsampling.cpp:5-106) on the GPU.

    grid_subsampling(points, features=None, labels=None, sampleDl=0.1, verbose=0)
        points (N, 3) float32, features (N, d) float32, labels (N,) or (N, L) int32
        -> sub_points [, sub_features] [, sub_labels]      (same return arity rules as the reference wrapper)

Inputs may be CUDA tensors, CPU tensors or numpy arrays (moved to the current CUDA device); outputs are CUDA tensors,
or numpy arrays when the points came in as numpy.  Barycentres and feature means are bit-exact with the reference;
cells come out in order of first appearance in the input and label ties go to the smallest label (the reference
leaves both to std::unordered_map's iteration order -- compare as sets)."""
import numpy as np
import torch

from . import _lib


def _dev(a, dtype):
    if a is None:
        return None
    t = torch.as_tensor(a)
    return t.to(device="cuda", dtype=dtype).contiguous()


def grid_subsampling(points, features=None, labels=None, sampleDl=0.1, verbose=0):
    as_numpy = isinstance(points, np.ndarray)
    pts = _dev(points, torch.float32)
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError("points must be (N, 3)")
    n = pts.shape[0]
    feats = _dev(features, torch.float32)
    labs = _dev(labels, torch.int32)
    lab_1d = labs is not None and labs.dim() == 1
    if lab_1d:
        labs = labs.reshape(n, 1)
    fdim = 0 if feats is None else feats.shape[1]
    ldim = 0 if labs is None else labs.shape[1]
    if (feats is not None and feats.shape[0] != n) or (labs is not None and labs.shape[0] != n):
        raise ValueError("features / labels must have one row per point")
    dev = pts.device
    lib = _lib.load()
    ws = torch.empty((max(lib.pasnl_grid_subsample_workspace_bytes(n), 4),), dtype=torch.uint8, device=dev)
    num = torch.zeros((1,), dtype=torch.int32, device=dev)
    sub_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    sub_f = torch.empty((n, fdim), dtype=torch.float32, device=dev) if fdim else None
    sub_l = torch.empty((n, ldim), dtype=torch.int32, device=dev) if ldim else None
    with torch.cuda.device(dev):
        _lib.call("pasnl_grid_subsample", n, _lib.ptr(pts), fdim, _lib.ptr(feats), ldim, _lib.ptr(labs), float(sampleDl),
                  _lib.ptr(ws), _lib.ptr(num), _lib.ptr(sub_p), _lib.ptr(sub_f), _lib.ptr(sub_l),
                  torch.cuda.current_stream().cuda_stream)
    m = int(num.item())                      # the output size is data dependent: one host read, like the reference's return
    out = [sub_p[:m]]
    if fdim:
        out.append(sub_f[:m])
    if ldim:
        out.append(sub_l[:m, 0] if lab_1d else sub_l[:m])
    if as_numpy:
        out = [o.cpu().numpy() for o in out]
    return out[0] if len(out) == 1 else tuple(out)
