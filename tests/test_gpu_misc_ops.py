"""prob_sample and select_top_k (csrc/misc_ops.cu) three-way: CUDA kernel vs the C oracle vs the reference's own kernels
compiled untouched (oracle/_ref/ref_sampling.so, ref_grouping.so).  Index-valued ops: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import cref, refnative

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,n,m", [(3, 5, 40), (4, 1000, 512), (2, 8192, 1000), (2, 8193, 1000), (3, 20001, 4096), (40, 77, 9)])
def test_prob_sample_matches_oracle_and_reference_kernel(cuda, b, n, m):
    from pointasnl_b200 import tf_sampling
    rng = np.random.default_rng(n + m)
    p = rng.uniform(0, 1, (b, n)).astype(np.float32)
    p[0, : n // 2] = 0.0                                   # zero-probability categories are never drawn
    r = rng.uniform(0, 1, (b, m)).astype(np.float32)
    got = tf_sampling.prob_sample(torch.from_numpy(p).to(cuda), torch.from_numpy(r).to(cuda))
    want, cum = cref.prob_sample(p, r)
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want)
    assert (p[0, got[0].cpu().numpy()] > 0).all()
    if refnative.available("ref_sampling.so"):
        ref_idx, ref_cum = refnative.gpu_prob_sample(torch.from_numpy(p).to(cuda), torch.from_numpy(r).to(cuda))
        assert np.array_equal(ref_cum.cpu().numpy(), cum)                     # the running sums, bit for bit
        assert np.array_equal(ref_idx.cpu().numpy(), want)


@pytest.mark.parametrize("b,m,n,k", [(2, 64, 128, 16), (1, 7, 33, 33), (3, 100, 50, 5), (2, 16, 300, 32)])
def test_select_top_k_matches_oracle_and_reference_kernel_including_ties(cuda, b, m, n, k):
    from pointasnl_b200 import tf_grouping
    rng = np.random.default_rng(k)
    d = rng.integers(0, 12, (b, m, n)).astype(np.float32)              # masses of exact ties
    d[0, 0] = 3.0
    idx, val = tf_grouping.select_top_k(k, torch.from_numpy(d).to(cuda))
    wi, wv = cref.selection_sort(k, d)
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(val.cpu().numpy(), wv)
    assert (np.diff(val.cpu().numpy()[..., :k], axis=-1) >= 0).all()
    if refnative.available("ref_grouping.so"):
        ri, rv = refnative.gpu_selection_sort(k, torch.from_numpy(d).to(cuda))
        assert np.array_equal(ri.cpu().numpy(), wi) and np.array_equal(rv.cpu().numpy(), wv)
