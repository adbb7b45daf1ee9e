"""PyTorch-CPU restatement of the reference's TF graph code on the hot path (TEST INFRASTRUCTURE, see
oracle/__init__.py).  **Parity unpinned**: TensorFlow 1.13 cannot be installed in this image, so these functions
follow utils/pointasnl_util.py line by line but cannot be executed against it.

Core-level functions (the maths the CUDA kernels fuse):
    as_core(q, kv, scale)                 pointasnl_util.py:137-146
    as_resample(logits, xyz, feat)        pointasnl_util.py:155,168-171
    nl_core(q, kv, scale)                 pointasnl_util.py:195-212 (mode='dot')
Layer / model level restatements live in oracle/model_ref.py.
"""
import torch


def as_core(q, kv, scale):
    """q (..., k, d); kv (..., k, 2d) = conv_kv_ds output: first d channels = keys, last d = values
    (pointasnl_util.py:137-138).  weights = softmax(q @ k^T * scale, -1) (:140-143); out = weights @ v (:146)."""
    d = q.shape[-1]
    kf, v = kv[..., :d], kv[..., d:]
    w = torch.matmul(q, kf.transpose(-1, -2)) * scale
    w = torch.softmax(w, dim=-1)
    return torch.matmul(w, v), w


def as_resample(logits, xyz, feat):
    """logits (b,s,k,1+cg) -> softmax over the k axis (:155, axis=2); new_xyz = sum_k w[...,0:1]*xyz (:168,170);
    new_feature = sum_k w[...,1:]*feat (:169,171)."""
    w = torch.softmax(logits, dim=2)
    new_xyz = (xyz * w[..., 0:1]).sum(dim=2)
    new_feat = (feat * w[..., 1:]).sum(dim=2)
    return new_xyz, new_feat, w


def nl_core(q, kv, scale):
    """q (b,s,d); kv (b,n,2d) = conv_kv output, keys then values (:195-196).
    attention_map = softmax(q @ k^T * scale, -1) (:198-201,211); out = attention_map @ v (:212)."""
    d = q.shape[-1]
    kf, v = kv[..., :d], kv[..., d:]
    a = torch.matmul(q, kf.transpose(-1, -2)) * scale
    a = torch.softmax(a, dim=-1)
    return torch.matmul(a, v)
