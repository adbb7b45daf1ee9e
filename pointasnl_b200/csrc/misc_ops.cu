// misc_ops.cu -- the reference ops that no shipped model calls but the op libraries export:
//   prob_sample      tf_ops/sampling/tf_sampling.py:13-22 -> cumsumKernel + binarysearchKernel (tf_sampling_g.cu:7-104)
//   selection sort   tf_ops/grouping/tf_grouping.py:22-32 -> selection_sort_gpu (tf_grouping_g.cu:83-123)
// Both are index-valued, so they are reproduced BIT-EXACTLY: prob_sample's answer depends on the rounding of an fp32
// running sum, hence the summation tree below is the reference's (groups of four, a two-sweep scan of the group totals,
// a compensated carry between 8192-element blocks); the selection sort keeps the reference's "first strict minimum,
// then swap" rule, which also fixes where equal distances and the unsorted remainder end up.
#include "common.cuh"

using namespace pasnl;

namespace mops {

constexpr int kBlockElems = 8192;            // elements per carry block (the reference's BlockSize * 4)
constexpr int kGroups = kBlockElems / 4;

// Inclusive prefix sum of row i of inp (b, n) into out (b, n).  One CTA per row.
__global__ void __launch_bounds__(512) cumsum_rows_kernel(int b, int n, const float *__restrict__ inp, float *__restrict__ out) {
    __shared__ float pre[kBlockElems];       // prefix inside each group of four
    __shared__ float tot[kGroups];           // group totals -> scanned in place
    for (int row = blockIdx.x; row < b; row += gridDim.x) {
        const float *src = inp + (size_t)row * n;
        float *dst = out + (size_t)row * n;
        float carry = 0.f, comp = 0.f;       // compensated running sum across blocks
        for (int j = 0; j < n; j += kBlockElems) {
            const int cnt = min(n - j, kBlockElems);
            const int padded = (cnt + 3) & ~3, groups = padded >> 2;
            for (int g = threadIdx.x; g < groups; g += blockDim.x) {
                const int k = g * 4;
                if (k + 3 < cnt) {
                    const float v1 = src[j + k], v2 = src[j + k + 1], v3 = src[j + k + 2], v4 = src[j + k + 3];
                    const float s12 = __fadd_rn(v2, v1), s34 = __fadd_rn(v4, v3);
                    pre[k] = v1;
                    pre[k + 1] = s12;
                    pre[k + 2] = __fadd_rn(v3, s12);
                    pre[k + 3] = __fadd_rn(s34, s12);
                    tot[g] = pre[k + 3];
                } else {                      // ragged last group: plain running sum, padding repeats the total
                    float v = 0.f;
                    for (int e = k; e < cnt; ++e) { v = __fadd_rn(v, src[j + e]); pre[e] = v; }
                    for (int e = cnt; e < padded; ++e) pre[e] = v;
                    tot[g] = v;
                }
            }
            // up-sweep: tot[(2q+2)<<u - 1] += tot[(2q+1)<<u - 1]
            int u = 0;
            for (; (2 << u) <= groups; ++u) {
                __syncthreads();
                for (int q = threadIdx.x; q < (groups >> (u + 1)); q += blockDim.x)
                    tot[(((q << 1) + 2) << u) - 1] = __fadd_rn(tot[(((q << 1) + 2) << u) - 1], tot[(((q << 1) + 1) << u) - 1]);
            }
            // down-sweep: tot[(2q+3)<<u - 1] += tot[(2q+2)<<u - 1]
            for (--u; u >= 0; --u) {
                __syncthreads();
                for (int q = threadIdx.x; q < ((groups - (1 << u)) >> (u + 1)); q += blockDim.x)
                    tot[(((q << 1) + 3) << u) - 1] = __fadd_rn(tot[(((q << 1) + 3) << u) - 1], tot[(((q << 1) + 2) << u) - 1]);
            }
            __syncthreads();
            for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
                const int g = e >> 2;
                const float in_block = g > 0 ? __fadd_rn(pre[e], tot[g - 1]) : pre[e];
                dst[j + e] = __fadd_rn(in_block, carry);
            }
            const float t = __fadd_rn(tot[groups - 1], comp);
            const float r2 = __fadd_rn(carry, t);
            comp = __fsub_rn(t, __fsub_rn(r2, carry));
            carry = r2;
            __syncthreads();
        }
    }
}

// out[i][j] = first index r with cum[i][r] >= q, q = query[i][j] * cum[i][n-1], found by the reference's descent
__global__ void binary_search_kernel(int b, int n, int m, const float *__restrict__ cum, const float *__restrict__ query,
                                     int *__restrict__ out) {
    int base = 1;
    while (base < n) base <<= 1;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)b * m) return;
    const int i = (int)(e / m);
    const float *row = cum + (size_t)i * n;
    const float q = __fmul_rn(__ldg(query + e), row[n - 1]);
    int r = n - 1;
    for (int k = base; k >= 1; k >>= 1)
        if (r >= k && row[r - k] >= q) r -= k;
    out[e] = r;
}

// One warp per (batch, row): k rounds of "find the first strict minimum of out[s..n), swap it to position s".
__global__ void __launch_bounds__(256) selection_sort_kernel(long nrows, int n, int k, const float *__restrict__ dist,
                                                             int *__restrict__ outi, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const float *src = dist + row * n;
    float *pd = out + row * n;
    int *pi = outi + row * n;
    for (int s = lane; s < n; s += 32) { pd[s] = src[s]; pi[s] = s; }
    __syncwarp();
    for (int s = 0; s < k && s < n; ++s) {
        float best = 3.4e38f * 3.4e38f;      // +inf
        int bidx = 0x7fffffff;
        for (int t = s + lane; t < n; t += 32) {
            const float v = pd[t];
            if (v < best) { best = v; bidx = t; }      // per lane: earliest index among its equal values
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(FULL_MASK, best, o);
            const int oi = __shfl_xor_sync(FULL_MASK, bidx, o);
            if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        // the reference starts from min = s and moves only on a STRICT '<': among equal minima the earliest index wins,
        // which is what (value, index) lexicographic arg-min yields; NaNs never compare less and stay where they are
        if (bidx == 0x7fffffff) bidx = s;
        if (lane == 0 && bidx != s) {
            const float tv = pd[bidx]; pd[bidx] = pd[s]; pd[s] = tv;
            const int ti = pi[bidx]; pi[bidx] = pi[s]; pi[s] = ti;
        }
        __syncwarp();
    }
}

}  // namespace mops

// Replaces probsampleLauncher(int b,int n,int m,const float* inp_p,const float* inp_r,float* temp,int* out)
// (tf_sampling_g.cu:194-197): temp (b,n) receives the running sums, out (b,m) the sampled indices.
extern "C" int pasnl_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out,
                                 void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 1 && m >= 0);
    if (b == 0 || m == 0) return PASNL_OK;
    PASNL_CHECK_ARG(inp_p && inp_r && temp && out);
    cudaStream_t st = (cudaStream_t)stream;
    mops::cumsum_rows_kernel<<<b < 1184 ? b : 1184, 512, 0, st>>>(b, n, inp_p, temp);
    const long tot = (long)b * m;
    mops::binary_search_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(b, n, m, temp, inp_r, out);
    return pasnl_launch_status();
}

// Replaces selectionSortLauncher(int b,int n,int m,int k,const float* dist,int* outi,float* out) (tf_grouping_g.cu:129-132).
extern "C" int pasnl_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 1 && m >= 0 && k >= 1);
    if (b == 0 || m == 0) return PASNL_OK;
    PASNL_CHECK_ARG(dist && outi && out);
    const long rows = (long)b * m;
    mops::selection_sort_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(rows, n, k, dist, outi, out);
    return pasnl_launch_status();
}
