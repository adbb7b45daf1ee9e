// small_ops.cu -- the launch-bound glue of a training step, as library kernels ("dense glue off
// cuBLAS / cuDNN", finished): the classifier head's skinny GEMMs (rows = batch size), residual adds, channel concats,
// dropout, the soft-max cross-entropy losses (value and gradient in one pass), and the L2 regulariser + Adam update
// over ONE flat parameter buffer.  None of these is bandwidth- or tensor-bound at the reference's sizes; what they buy
// is (a) no cuBLAS / ATen kernel on the step and (b) one launch where eager PyTorch issues three to sixty.
//
// Reference semantics: tf_util.fully_connected (utils/tf_util.py:318-355), tf.add / tf.concat in the layer functions
// (utils/pointasnl_util.py:283-285, utils/pointnet_util.py:218), tf_util.dropout (:590-615, tf.nn.dropout: keep with
// probability keep_prob, scale kept values by 1/keep_prob), tf.nn.sparse_softmax_cross_entropy_with_logits +
// reduce_mean (models/pointasnl_cls.py:60-61), tf.losses.sparse_softmax_cross_entropy with weights
// (SUM_BY_NONZERO_WEIGHTS, models/pointasnl_sem_seg_res.py:81), tf.nn.l2_loss over the 'weights' variables (:79) and
// tf.train.AdamOptimizer (train.py:153: beta1 0.9, beta2 0.999, epsilon 1e-8, lr_t = lr*sqrt(1-b2^t)/(1-b1^t)).
#include "common.cuh"

using namespace pasnl;

namespace sops {

// ------------------------------------------------------------------------------------------------ skinny GEMM
// out[r][j] (+)= sum_i in[r][i] * M(i, j),  r < rows <= RMAX.  M(i, j) = W[i*ldw + j] (TRANS = 0: y = x.W with W stored
// [k][n]) or W[j*ldw + i] (TRANS = 1: dx = dy.W^T with the same stored W).  CTA = 64 output columns x one 128-wide slab
// of the contraction; 4 thread groups split the slab; partial sums meet in shared memory and leave with atomics.
constexpr int kSkJ = 64, kSkI = 128, kSkThreads = 256;

template <int RMAX, bool TRANS>
__global__ void __launch_bounds__(kSkThreads) gemm_skinny_kernel(int rows, int ni, int nj, const float *__restrict__ in,
                                                                 long ldin, const float *__restrict__ W, long ldw,
                                                                 float *__restrict__ out, long ldout) {
    extern __shared__ __align__(16) float sm[];
    float *in_s = sm;                                   // [RMAX][kSkI]
    float *m_s = sm + RMAX * kSkI;                      // TRANS: [kSkJ][kSkI + 1]; reused as the reduction buffer
    const int t = threadIdx.x, jl = t & 63, iq = t >> 6;
    const int j0 = blockIdx.x * kSkJ, i0 = blockIdx.y * kSkI;
    for (int e = t; e < RMAX * kSkI; e += kSkThreads) {
        const int r = e / kSkI, ii = e - r * kSkI;
        in_s[e] = (r < rows && i0 + ii < ni) ? __ldg(in + (long)r * ldin + i0 + ii) : 0.f;
    }
    if (TRANS) {
        for (int e = t; e < kSkJ * kSkI; e += kSkThreads) {
            const int jj = e / kSkI, ii = e - jj * kSkI;
            m_s[jj * (kSkI + 1) + ii] = (j0 + jj < nj && i0 + ii < ni) ? __ldg(W + (long)(j0 + jj) * ldw + i0 + ii) : 0.f;
        }
    }
    __syncthreads();
    float acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r] = 0.f;
    const bool jv = j0 + jl < nj;
    for (int ii = iq * 32; ii < iq * 32 + 32; ii += 4) {
        float w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (TRANS) w[u] = m_s[jl * (kSkI + 1) + ii + u];
            else w[u] = (jv && i0 + ii + u < ni) ? __ldg(W + (long)(i0 + ii + u) * ldw + j0 + jl) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const float4 x = *reinterpret_cast<const float4 *>(in_s + r * kSkI + ii);
            acc[r] = fmaf(x.x, w[0], fmaf(x.y, w[1], fmaf(x.z, w[2], fmaf(x.w, w[3], acc[r]))));
        }
    }
    __syncthreads();                                     // m_s / in_s are dead: reuse as [4][RMAX][64]
    float *red = sm;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) red[(iq * RMAX + r) * kSkJ + jl] = acc[r];
    __syncthreads();
    for (int e = t; e < RMAX * kSkJ; e += kSkThreads) {
        const int r = e / kSkJ, j = e - r * kSkJ;
        if (r < rows && j0 + j < nj) {
            const float s = (red[(0 * RMAX + r) * kSkJ + j] + red[(1 * RMAX + r) * kSkJ + j]) +
                            (red[(2 * RMAX + r) * kSkJ + j] + red[(3 * RMAX + r) * kSkJ + j]);
            atomicAdd(out + (long)r * ldout + j0 + j, s);
        }
    }
}

// dW[i][j] = sum_r x[r][i] * dy[r][j], rows <= 64: 64 x 64 output tile per CTA, 4 x 4 per thread, written (not added)
__global__ void __launch_bounds__(256) wgrad_skinny_kernel(int rows, int ni, int nj, const float *__restrict__ x, long ldx,
                                                           const float *__restrict__ dy, long ldy,
                                                           float *__restrict__ dW, long ldw) {
    __shared__ __align__(16) float xs[64][64], ds[64][64];
    const int t = threadIdx.x, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    for (int e = t; e < 64 * 64; e += 256) {
        const int r = e >> 6, c = e & 63;
        xs[r][c] = (r < rows && i0 + c < ni) ? __ldg(x + (long)r * ldx + i0 + c) : 0.f;
        ds[r][c] = (r < rows && j0 + c < nj) ? __ldg(dy + (long)r * ldy + j0 + c) : 0.f;
    }
    __syncthreads();
    const int ti = (t >> 4) * 4, tj = (t & 15) * 4;
    float acc[4][4] = {};
    for (int r = 0; r < rows; ++r) {
        const float4 a = *reinterpret_cast<const float4 *>(&xs[r][ti]);
        const float4 b = *reinterpret_cast<const float4 *>(&ds[r][tj]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (i0 + ti + u < ni && j0 + tj + v < nj) dW[(long)(i0 + ti + u) * ldw + j0 + tj + v] = acc[u][v];
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void add_n_kernel(long n, const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ c,
                             float *__restrict__ out, int vec) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        if (i * 4 + 3 < n) {
            float4 x = __ldg(reinterpret_cast<const float4 *>(a) + i), y = __ldg(reinterpret_cast<const float4 *>(b) + i);
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            if (c) {
                const float4 z = __ldg(reinterpret_cast<const float4 *>(c) + i);
                x.x += z.x; x.y += z.y; x.z += z.z; x.w += z.w;
            }
            reinterpret_cast<float4 *>(out)[i] = x;
        } else {
            for (long e = i * 4; e < n; ++e) out[e] = (a[e] + b[e]) + (c ? c[e] : 0.f);
        }
    } else if (i < n) {
        out[i] = (a[i] + b[i]) + (c ? c[i] : 0.f);
    }
}

// out[r] = [a[r] | b[r]] (forward) ; ga[r], gb[r] = the two column ranges of g[r] (backward)
__global__ void concat_cols_kernel(long rows, int ca, int cb, const float *__restrict__ a, long lda,
                                   const float *__restrict__ b, long ldb, float *__restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = ca + cb;
    if (e >= rows * w) return;
    const long r = e / w;
    const int c = (int)(e - r * w);
    out[e] = c < ca ? __ldg(a + r * lda + c) : __ldg(b + r * ldb + (c - ca));
}
__global__ void split_cols_kernel(long rows, int ca, int cb, const float *__restrict__ g, float *__restrict__ ga,
                                  float *__restrict__ gb) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = ca + cb;
    if (e >= rows * w) return;
    const long r = e / w;
    const int c = (int)(e - r * w);
    const float v = __ldg(g + e);
    if (c < ca) { if (ga) ga[r * ca + c] = v; }
    else if (gb) gb[r * cb + (c - ca)] = v;
}

// counter-based uniform in [0,1): 2 rounds of a 64-bit mix of (seed, step, element index)
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t step, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (step + 1) + idx * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}
// out = x * mask / keep_prob with mask ~ Bernoulli(keep_prob) per element; the mask is a pure function of
// (seed, *step, index), so the backward pass regenerates it instead of storing it.
__global__ void dropout_kernel(long n, const float *__restrict__ x, float keep_prob, uint64_t seed,
                               const long *__restrict__ step, float *__restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t st = step ? (uint64_t)__ldg(step) : 0ull;
    out[i] = uniform01(seed, st, (uint64_t)i) < keep_prob ? x[i] / keep_prob : 0.f;
}

// ------------------------------------------------------------------------------------------------ soft-max cross entropy
// One warp per row.  fwd: ws[0] += w_r * ce_r, ws[1] += (w_r != 0); lse[r] saved.  The LAST CTA to finish publishes
// loss = out_scale * ws[0] / denom  (denom = rows when weights == NULL: reduce_mean; else max(ws[1], 1)) and leaves
// denom in ws[2] for the backward.  bwd: dlogits[r][c] = g * out_scale * w_r / denom * (softmax - onehot).
__global__ void __launch_bounds__(256) softmax_ce_fwd_kernel(long rows, int c, const float *__restrict__ logits,
                                                             const long *__restrict__ labels,
                                                             const float *__restrict__ weights, float out_scale,
                                                             float *__restrict__ lse, double *__restrict__ ws,
                                                             unsigned *__restrict__ done, float *__restrict__ loss) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long r = (long)blockIdx.x * 8 + warp;
    double part = 0.0, nz = 0.0;
    if (r < rows) {
        const float *p = logits + r * c;
        float mx = -3.4e38f;
        for (int j = lane; j < c; j += 32) mx = fmaxf(mx, __ldg(p + j));
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
        float se = 0.f;
        for (int j = lane; j < c; j += 32) se += expf(__ldg(p + j) - mx);
#pragma unroll
        for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(FULL_MASK, se, o);
        const float l = mx + logf(se);
        if (lane == 0) {
            lse[r] = l;
            const float w = weights ? __ldg(weights + r) : 1.f;
            const long lab = __ldg(labels + r);
            part = (double)w * (double)(l - __ldg(p + lab));
            nz = (w != 0.f) ? 1.0 : 0.0;
        }
    }
    __shared__ double sp[8], sn[8];
    if (lane == 0) { sp[warp] = part; sn[warp] = nz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 8; ++i) { a += sp[i]; b += sn[i]; }
        atomicAdd(ws, a);
        atomicAdd(ws + 1, b);
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            const double tot = atomicAdd(ws, 0.0), cnt = atomicAdd(ws + 1, 0.0);
            const double denom = weights ? (cnt > 1.0 ? cnt : 1.0) : (double)rows;
            ws[2] = denom;
            *loss = (float)((double)out_scale * tot / denom);
        }
    }
}
__global__ void __launch_bounds__(256) softmax_ce_bwd_kernel(long rows, int c, const float *__restrict__ logits,
                                                             const long *__restrict__ labels,
                                                             const float *__restrict__ weights, float out_scale,
                                                             const float *__restrict__ lse, const double *__restrict__ ws,
                                                             const float *__restrict__ gout, float *__restrict__ dlogits) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * c) return;
    const long r = e / c;
    const int j = (int)(e - r * c);
    const float w = weights ? __ldg(weights + r) : 1.f;
    const float g = __ldg(gout) * out_scale * w / (float)ws[2];
    const float p = expf(__ldg(logits + e) - __ldg(lse + r));
    dlogits[e] = g * (p - (j == (int)__ldg(labels + r) ? 1.f : 0.f));
}

// ------------------------------------------------------------------------------------------------ L2 + Adam, flat
// p, g, m, v: one flat fp32 buffer each.  l2_mask[i] != 0 marks the elements of the 'weights' variables: they get the
// regulariser gradient l2_coeff * p (d/dp of l2_coeff * sum(p^2)/2), and l2_coeff * p^2 / 2 is accumulated into *reg
// (computed from the PRE-update value, i.e. the loss of the step that produced g).  grad_scale multiplies g first (the
// 1/world of a summed all-reduce).  *step is incremented by thread 0 of block 0 AFTER the grid has read it... it is
// read by every thread at entry and bumped by the host-side launcher's second tiny kernel to stay race-free.
__global__ void adam_l2_kernel(long n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                               float *__restrict__ v, const unsigned char *__restrict__ l2_mask, float l2_coeff,
                               float grad_scale, float lr, float b1, float b2, float eps, const long *__restrict__ step,
                               double *__restrict__ reg) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float s_lr_t;
    if (threadIdx.x == 0) {               // lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), once per block
        const double t = (double)(__ldg(step) + 1);
        s_lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
    }
    __syncthreads();
    const float lr_t = s_lr_t;
    float r = 0.f;
    if (i < n) {
        const float pi = p[i];
        float gi = g[i] * grad_scale;
        if (l2_mask && l2_mask[i]) {
            gi = fmaf(l2_coeff, pi, gi);
            r = 0.5f * l2_coeff * pi * pi;
        }
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        // tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)
        p[i] = pi - lr_t * mi / (sqrtf(vi) + eps);
    }
    if (reg) {
#pragma unroll
        for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(FULL_MASK, r, o);
        __shared__ float sr[32];
        if ((threadIdx.x & 31) == 0) sr[threadIdx.x >> 5] = r;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += (double)sr[w];
            if (a != 0.0) atomicAdd(reg, a);
        }
    }
}
// after the update: step += 1; total loss = data loss + regulariser; the regulariser accumulator is cleared for the next step
__global__ void bump_step_kernel(long *step, double *reg, const float *loss_in, float *loss_out) {
    *step += 1;
    if (loss_out) *loss_out = (loss_in ? *loss_in : 0.f) + (reg ? (float)*reg : 0.f);
    if (reg) *reg = 0.0;
}

// ------------------------------------------------------------------------------------------------ SA-layer front end
// out[g, c] = max_t x[g, t, c] for k <= 64 neighbours, plus a 64-bit mask of the positions that attain the maximum
// (ties included: TF's reduce_max gradient is split evenly among them).  With the mask the backward pass is a SCATTER
// of groups*c values into an existing gradient buffer instead of a dense (groups, k, c) tensor that autograd would
// then have to add to the other gradients of the same activation (three 140 MB passes on cls layer 2).
__global__ void __launch_bounds__(256) maxpool_k_mask_kernel(long gc, int k, int c, const float *__restrict__ x,
                                                             float *__restrict__ out,
                                                             unsigned long long *__restrict__ mask) {
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < gc; e += (long)gridDim.x * 256) {
        const long g = e / c;
        const int ch = (int)(e - g * c);
        const float *p = x + g * (long)k * c + ch;
        float m = __ldg(p);
        unsigned long long bits = 1ull;
        for (int t = 1; t < k; ++t) {
            const float v = __ldg(p + (long)t * c);
            if (v > m) { m = v; bits = 1ull << t; }
            else if (v == m) bits |= 1ull << t;
        }
        out[e] = m;
        mask[e] = bits;
    }
}
// grad_x[g, t, c] += grad_out[g, c] / popcount(mask[g, c]) for every t in the mask (in place)
__global__ void __launch_bounds__(256) maxpool_k_scatter_kernel(long gc, int k, int c,
                                                                const unsigned long long *__restrict__ mask,
                                                                const float *__restrict__ grad_out,
                                                                float *__restrict__ grad_x) {
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < gc; e += (long)gridDim.x * 256) {
        const long g = e / c;
        const int ch = (int)(e - g * c);
        unsigned long long bits = __ldg(mask + e);
        const float share = __ldg(grad_out + e) / (float)__popcll(bits);
        float *q = grad_x + g * (long)k * c + ch;
        while (bits) {
            const int t = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            q[(long)t * c] += share;
        }
    }
}
// dst[r][0..nc) = src[r][0..nc) (src row pitch lds): the leading-column slice as a dense tensor
__global__ void take_cols_kernel(long rows, int nc, const float *__restrict__ src, long lds, float *__restrict__ dst) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * nc) return;
    const long r = e / nc;
    const int c = (int)(e - r * nc);
    dst[e] = __ldg(src + r * lds + c);
}
// dst[r][0..nc) += src[r][0..nc)  (dst row pitch ldd): the gradient of a leading-column slice, added in place
__global__ void add_cols_kernel(long rows, int nc, const float *__restrict__ src, float *__restrict__ dst, long ldd) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * nc) return;
    const long r = e / nc;
    const int c = (int)(e - r * nc);
    dst[r * ldd + c] += __ldg(src + e);
}

// ------------------------------------------------------------------------------------------------ weight transposes
// table[4*i .. 4*i+3] = {src pointer, dst pointer, rows, cols} of matrix i: dst (cols, rows) = src (rows, cols)^T.
// One launch refreshes the K-major copies of EVERY layer's weights after an optimizer step (blockIdx.y = matrix).
__global__ void __launch_bounds__(256) transpose_batched_kernel(const long *__restrict__ table) {
    __shared__ float tile[32][33];
    const long *e = table + 4 * (long)blockIdx.y;
    const float *src = reinterpret_cast<const float *>(e[0]);
    float *dst = reinterpret_cast<float *>(e[1]);
    const int rows = (int)e[2], cols = (int)e[3];
    const int tc = (cols + 31) / 32, tr = (rows + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int tI = blockIdx.x; tI < tc * tr; tI += gridDim.x) {
        const int r0 = (tI / tc) * 32, c0 = (tI % tc) * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + ty + 8 * j, cc = c0 + tx;
            tile[ty + 8 * j][tx] = (r < rows && cc < cols) ? __ldg(src + (long)r * cols + cc) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cc = c0 + ty + 8 * j, r = r0 + tx;
            if (cc < cols && r < rows) dst[(long)cc * rows + r] = tile[tx][ty + 8 * j];
        }
        __syncthreads();
    }
}

// out[r][c] = (x ? x[r][c] : 0) + bias[c]   (tf.nn.bias_add; also the bias-initialised accumulator of a skinny GEMM)
__global__ void bias_add_kernel(long rows, int n, const float *__restrict__ x, const float *__restrict__ bias,
                                float *__restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * n) return;
    const int c = (int)(e % n);
    out[e] = (x ? __ldg(x + e) : 0.f) + __ldg(bias + c);
}
// out[c] += sum_r x[r][c]  (the bias gradient); out zero-initialised by the launcher
__global__ void __launch_bounds__(256) col_sum_kernel(long rows, int n, const float *__restrict__ x, float *__restrict__ out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    const long chunk = (rows + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
    float s = 0.f;
    if (c < n)
        for (long r = r0 + ry; r < r1; r += 8) s += __ldg(x + r * n + c);
    __shared__ float red[8][33];
    red[ry][threadIdx.x & 31] = s;
    __syncthreads();
    if (ry == 0 && c < n) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
        atomicAdd(out + c, t);
    }
}

__global__ void bn_finalize_kernel(long rows, int c, const double *__restrict__ ws, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, float eps, const float *__restrict__ bias,
                                   float *__restrict__ running_mean, float *__restrict__ running_var, float decay,
                                   float vcorr, float *__restrict__ mean, float *__restrict__ var,
                                   float *__restrict__ scale, float *__restrict__ shift) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const double n = (double)rows, m = ws[ch] / n;
    double vv = ws[c + ch] / n - m * m;
    if (vv < 0.0) vv = 0.0;
    const float mf = (float)m, vf = (float)vv;
    const float sc = rsqrtf(vf + eps) * __ldg(gamma + ch);
    mean[ch] = mf;
    var[ch] = vf;
    scale[ch] = sc;
    shift[ch] = fmaf(-mf, sc, __ldg(beta + ch));
    if (running_mean) running_mean[ch] = decay * running_mean[ch] + (1.f - decay) * (mf + (bias ? bias[ch] : 0.f));
    if (running_var) running_var[ch] = decay * running_var[ch] + (1.f - decay) * vf * vcorr;
}

}  // namespace sops

extern "C" int pasnl_gemm_skinny(int rows, int ni, int nj, const float *in, long ldin, const float *W, long ldw,
                                 int w_transposed, float *out, long ldout, void *stream) {
    PASNL_CHECK_ARG(rows >= 1 && rows <= 64 && ni >= 1 && nj >= 1 && in && W && out && ldin >= ni && ldout >= nj);
    PASNL_CHECK_ARG(w_transposed ? ldw >= ni : ldw >= nj);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((nj + sops::kSkJ - 1) / sops::kSkJ, (ni + sops::kSkI - 1) / sops::kSkI);
    const int rmax = rows <= 32 ? 32 : 64;
    const size_t fl = (size_t)rmax * sops::kSkI + (w_transposed ? (size_t)sops::kSkJ * (sops::kSkI + 1) : 0);
    const size_t red = (size_t)4 * rmax * sops::kSkJ;
    const size_t smem = sizeof(float) * (fl > red ? fl : red);
#define PASNL_SK(R, T)                                                                                                  \
    do {                                                                                                                \
        auto kern = sops::gemm_skinny_kernel<R, T>;                                                                     \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
        if (e != cudaSuccess) return (int)e;                                                                            \
        kern<<<grid, sops::kSkThreads, smem, st>>>(rows, ni, nj, in, ldin, W, ldw, out, ldout);                         \
    } while (0)
    if (rmax == 32) { if (w_transposed) PASNL_SK(32, true); else PASNL_SK(32, false); }
    else            { if (w_transposed) PASNL_SK(64, true); else PASNL_SK(64, false); }
#undef PASNL_SK
    return pasnl_launch_status();
}

extern "C" int pasnl_wgrad_skinny(int rows, int ni, int nj, const float *x, long ldx, const float *dy, long ldy, float *dW,
                                  long ldw, void *stream) {
    PASNL_CHECK_ARG(rows >= 1 && rows <= 64 && ni >= 1 && nj >= 1 && x && dy && dW && ldx >= ni && ldy >= nj && ldw >= nj);
    const dim3 grid((nj + 63) / 64, (ni + 63) / 64);
    sops::wgrad_skinny_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, ni, nj, x, ldx, dy, ldy, dW, ldw);
    return pasnl_launch_status();
}

extern "C" int pasnl_add_n(long n, const float *a, const float *b, const float *c, float *out, void *stream) {
    PASNL_CHECK_ARG(n >= 0 && (n == 0 || (a && b && out)));
    if (n == 0) return PASNL_OK;
    auto al = [](const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int vec = al(a) && al(b) && al(c) && al(out);
    const long work = vec ? (n + 3) / 4 : n;
    sops::add_n_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, a, b, c, out, vec);
    return pasnl_launch_status();
}

extern "C" int pasnl_concat_cols(long rows, int ca, int cb, const float *a, long lda, const float *b, long ldb,
                                 float *out, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && ca >= 1 && cb >= 1 && lda >= ca && ldb >= cb && (rows == 0 || (a && b && out)));
    if (rows == 0) return PASNL_OK;
    const long tot = rows * (ca + cb);
    sops::concat_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, ca, cb, a, lda, b, ldb, out);
    return pasnl_launch_status();
}

extern "C" int pasnl_split_cols(long rows, int ca, int cb, const float *g, float *ga, float *gb, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && ca >= 1 && cb >= 1 && (rows == 0 || g));
    if (rows == 0) return PASNL_OK;
    const long tot = rows * (ca + cb);
    sops::split_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, ca, cb, g, ga, gb);
    return pasnl_launch_status();
}

extern "C" int pasnl_dropout(long n, const float *x, float keep_prob, unsigned long long seed, const long *step,
                             float *out, void *stream) {
    PASNL_CHECK_ARG(n >= 0 && keep_prob > 0.f && keep_prob <= 1.f && (n == 0 || (x && out)));
    if (n == 0) return PASNL_OK;
    sops::dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, x, keep_prob, seed, step, out);
    return pasnl_launch_status();
}

extern "C" size_t pasnl_softmax_ce_workspace_bytes(void) { return 4 * sizeof(double); }

extern "C" int pasnl_softmax_ce_fwd(long rows, int c, const float *logits, const long *labels, const float *weights,
                                    float out_scale, float *lse, void *workspace, float *loss, void *stream) {
    PASNL_CHECK_ARG(rows >= 1 && c >= 1 && logits && labels && lse && workspace && loss);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, pasnl_softmax_ce_workspace_bytes(), st);
    if (e != cudaSuccess) return (int)e;
    double *ws = (double *)workspace;
    sops::softmax_ce_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(rows, c, logits, labels, weights, out_scale, lse,
                                                                          ws, (unsigned *)(ws + 3), loss);
    return pasnl_launch_status();
}

extern "C" int pasnl_softmax_ce_bwd(long rows, int c, const float *logits, const long *labels, const float *weights,
                                    float out_scale, const float *lse, const void *workspace, const float *grad_loss,
                                    float *dlogits, void *stream) {
    PASNL_CHECK_ARG(rows >= 1 && c >= 1 && logits && labels && lse && workspace && grad_loss && dlogits);
    const long tot = rows * c;
    sops::softmax_ce_bwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rows, c, logits, labels, weights, out_scale, lse, (const double *)workspace, grad_loss, dlogits);
    return pasnl_launch_status();
}

extern "C" int pasnl_adam_l2(long n, float *p, const float *g, float *m, float *v, const unsigned char *l2_mask,
                             float l2_coeff, float grad_scale, float lr, float beta1, float beta2, float eps, long *step,
                             double *reg_loss, const float *loss_in, float *loss_out, void *stream) {
    PASNL_CHECK_ARG(n >= 1 && p && g && m && v && step);
    cudaStream_t st = (cudaStream_t)stream;
    sops::adam_l2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, p, g, m, v, l2_mask, l2_coeff, grad_scale, lr, beta1,
                                                                      beta2, eps, step, reg_loss);
    sops::bump_step_kernel<<<1, 1, 0, st>>>(step, reg_loss, loss_in, loss_out);
    return pasnl_launch_status();
}

extern "C" int pasnl_maxpool_k_mask(long groups, int k, int c, const float *x, float *out, unsigned long long *mask,
                                    void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && k <= 64 && c >= 1 && (groups == 0 || (x && out && mask)));
    if (groups == 0) return PASNL_OK;
    long blocks = (groups * c + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    sops::maxpool_k_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(groups * c, k, c, x, out, mask);
    return pasnl_launch_status();
}

extern "C" int pasnl_maxpool_k_scatter(long groups, int k, int c, const unsigned long long *mask, const float *grad_out,
                                       float *grad_x, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && k <= 64 && c >= 1 && (groups == 0 || (mask && grad_out && grad_x)));
    if (groups == 0) return PASNL_OK;
    long blocks = (groups * c + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    sops::maxpool_k_scatter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(groups * c, k, c, mask, grad_out, grad_x);
    return pasnl_launch_status();
}

extern "C" int pasnl_add_cols(long rows, int nc, const float *src, float *dst, long ldd, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && nc >= 1 && ldd >= nc && (rows == 0 || (src && dst)));
    if (rows == 0) return PASNL_OK;
    const long tot = rows * nc;
    sops::add_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, nc, src, dst, ldd);
    return pasnl_launch_status();
}

extern "C" int pasnl_transpose_batched(int count, const long *table, int max_tiles, void *stream) {
    PASNL_CHECK_ARG(count >= 0 && (count == 0 || table) && max_tiles >= 1);
    if (count == 0) return PASNL_OK;
    const dim3 grid((unsigned)(max_tiles < 64 ? max_tiles : 64), (unsigned)count);
    sops::transpose_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table);
    return pasnl_launch_status();
}

extern "C" int pasnl_bias_add(long rows, int n, const float *x, const float *bias, float *out, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && n >= 1 && (rows == 0 || (bias && out)));
    if (rows == 0) return PASNL_OK;
    const long tot = rows * n;
    sops::bias_add_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, n, x, bias, out);
    return pasnl_launch_status();
}

extern "C" int pasnl_col_sum(long rows, int n, const float *x, float *out, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && n >= 1 && out && (rows == 0 || x));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n, st);
    if (e != cudaSuccess) return (int)e;
    if (rows == 0) return PASNL_OK;
    long slabs = rows / 2048;
    if (slabs < 1) slabs = 1;
    if (slabs > 256) slabs = 256;
    const dim3 grid((unsigned)((n + 31) / 32), (unsigned)slabs);
    sops::col_sum_kernel<<<grid, 256, 0, st>>>(rows, n, x, out);
    return pasnl_launch_status();
}

extern "C" int pasnl_take_cols(long rows, int nc, const float *src, long lds, float *dst, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && nc >= 1 && lds >= nc && (rows == 0 || (src && dst)));
    if (rows == 0) return PASNL_OK;
    const long tot = rows * nc;
    sops::take_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, nc, src, lds, dst);
    return pasnl_launch_status();
}

extern "C" int pasnl_bn_finalize(long rows, int c, const void *ws, const float *gamma, const float *beta, float eps,
                                 const float *bias, float *running_mean, float *running_var, float decay,
                                 int unbiased_moving_var, float *mean, float *var, float *scale, float *shift, void *stream) {
    PASNL_CHECK_ARG(rows >= 1 && c >= 1 && ws && gamma && beta && mean && var && scale && shift);
    const float vcorr = (unbiased_moving_var && rows > 1) ? (float)((double)rows / (double)(rows - 1)) : 1.f;
    sops::bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rows, c, (const double *)ws, gamma, beta, eps, bias,
                                                                               running_mean, running_var, decay, vcorr, mean,
                                                                               var, scale, shift);
    return pasnl_launch_status();
}
