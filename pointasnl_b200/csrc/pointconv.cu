// pointconv.cu -- the PointConv-style contraction of the SA / decoding layers (sm_100a).
//
// Replaces  tf.transpose(new_point, [0,1,3,2]) ; tf.matmul(new_point, weight)  (reference utils/pointasnl_util.py:272-273,
// 330-333): per neighbourhood g = (b, s) the K x C feature tile is contracted with the K x 32 weight-net output,
//     M[g, c, w] = sum_k X[g, k, c] * W[g, k, w].
// As batched (C x K)(K x 32) matmuls this is 4096..16384 tiny GEMMs (K = 32/64, C = 64/128): cuBLAS batched SGEMM plus a
// materialised transpose in PyTorch, ~0.95 ms per cls step forward+backward.  The tiles are far below a tcgen05 tile
// per group and the op is HBM bound (read X once: 4*G*K*(C+32) bytes, write 4*G*C*32), so: one CTA per neighbourhood,
// X and W staged in shared memory by coalesced float4 loads, fp32 FMA on 4x4 register tiles fed by LDS.128.
#include "common.cuh"

using namespace pasnl;

namespace {

constexpr int kT = 256;
constexpr int kWD = 32;   // weight_net_hidden(..., [32], ...): the reference hard-codes 32 (pointasnl_util.py:271)

__device__ __forceinline__ void load_tile(float *dst, const float *__restrict__ src, int nfloats) {
    for (int v = threadIdx.x; v < nfloats / 4; v += kT)
        reinterpret_cast<float4 *>(dst)[v] = __ldg(reinterpret_cast<const float4 *>(src) + v);
}

// Optional operand prologue (x_scale != NULL): the X tile that arrived in shared memory is the PRE-batch-norm output of the
// producing layer; it is replaced in place by act(x * scale[ch] + shift[ch]) (act 1 relu / 2 leaky_relu(0.2)) before the
// contraction, so that layer's normalised activation is never written to HBM (see pasnl_gemm_rows_ws's prologue).
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bn_act_tile(float *X, int nfloats, int c, const float *__restrict__ sc,
                                            const float *__restrict__ sh, int act) {
    for (int v = threadIdx.x; v < nfloats / 4; v += blockDim.x) {
        const int ch = (v * 4) % c;
        float4 q = reinterpret_cast<float4 *>(X)[v];
        const float4 a = __ldg(reinterpret_cast<const float4 *>(sc + ch)), b = __ldg(reinterpret_cast<const float4 *>(sh + ch));
        q.x = fmaf(q.x, a.x, b.x); q.y = fmaf(q.y, a.y, b.y); q.z = fmaf(q.z, a.z, b.z); q.w = fmaf(q.w, a.w, b.w);
        if (act == 1) {
            q.x = fmaxf(q.x, 0.f); q.y = fmaxf(q.y, 0.f); q.z = fmaxf(q.z, 0.f); q.w = fmaxf(q.w, 0.f);
        } else if (act == 2) {
            q.x = q.x > 0.f ? q.x : 0.2f * q.x; q.y = q.y > 0.f ? q.y : 0.2f * q.y;
            q.z = q.z > 0.f ? q.z : 0.2f * q.z; q.w = q.w > 0.f ? q.w : 0.2f * q.w;
        }
        reinterpret_cast<float4 *>(X)[v] = q;
    }
}

// Forward.  Thread tile = 4 channels x 4 weight columns: per neighbour t one LDS.128 of X[t][c..c+3] and one of
// W[t][w..w+3] feed 16 FMAs.  wg = t & 7 (8 groups of 4 weight columns), cg = t >> 3 (groups of 4 channels): a warp
// stores 4 full 128-byte rows of the (C x 32) output per instruction.
// The op streams (read X and W once, write M once), so the loads must never stop: a group's X and W tiles are
// contiguous in global memory and arrive by TMA bulk copies (cp.async.bulk + mbarrier) into a DOUBLE buffer -- the
// tiles of the CTA's next group are in flight while the current one is contracted.  (The first version loaded,
// synchronised, computed, synchronised per group: 43 % of the HBM peak.)
__global__ void __launch_bounds__(kT) pointconv_fwd_kernel(int groups, int k, int c, const float *__restrict__ x,
                                                           const float *__restrict__ w, float *__restrict__ out,
                                                           const float *__restrict__ x_scale,
                                                           const float *__restrict__ x_shift, int x_act) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar[2];
    const int xs = k * c, wsz = k * kWD;
    float *Xb[2] = {smem, smem + xs + wsz}, *Wb[2] = {smem + xs, smem + 2 * xs + wsz};
    const int wg = threadIdx.x & 7, cg0 = threadIdx.x >> 3;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int g, int b) {                          // one thread: both tiles of group g -> buffer b
        mbar_expect_tx(&bar[b], (uint32_t)(xs + wsz) * 4u);
        bulk_g2s(Xb[b], x + (size_t)g * xs, (uint32_t)xs * 4u, &bar[b]);
        bulk_g2s(Wb[b], w + (size_t)g * wsz, (uint32_t)wsz * 4u, &bar[b]);
    };
    if (threadIdx.x == 0 && (int)blockIdx.x < groups) issue(blockIdx.x, 0);
    int it = 0;
    for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
        const int b = it & 1, gn = g + gridDim.x;
        // buffer b^1 was last read in iteration it-1, which ended with a CTA barrier
        if (threadIdx.x == 0 && gn < groups) issue(gn, b ^ 1);
        mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));
        if (x_scale != nullptr) {
            bn_act_tile(Xb[b], xs, c, x_scale, x_shift, x_act);
            fence_async_smem();                 // the tile is rewritten by a later bulk copy (async proxy)
            __syncthreads();
        }
        const float *X = Xb[b], *W = Wb[b];
        float *o = out + (size_t)g * c * kWD;
        for (int cg = cg0; cg < c / 4; cg += blockDim.x / 8) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int t = 0; t < k; ++t) {
                const float4 xv = *reinterpret_cast<const float4 *>(X + t * c + cg * 4);
                const float4 wv = *reinterpret_cast<const float4 *>(W + t * kWD + wg * 4);
                const float xr[4] = {xv.x, xv.y, xv.z, xv.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], wr[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4 *>(o + (cg * 4 + i) * kWD + wg * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
    }
}

// Backward.  dM is staged twice: row-major [ch][32] (for dW) and transposed [w][C] (for dX), so both products read
// float4 operands.  dX tile = 4 t x 4 ch, dW tile = 4 t x 4 w; the two tile sets share ONE index space so that the
// threads left over by the dX tiles work on dW tiles in the same pass.  X, W and dM tiles arrive by TMA bulk copies;
// with NBUF = 2 (small tiles) the next group's tiles are in flight while this one is processed.
template <int NBUF>
__global__ void __launch_bounds__(kT) pointconv_bwd_kernel(int groups, int k, int c, const float *__restrict__ x,
                                                           const float *__restrict__ w,
                                                           const float *__restrict__ grad_out,
                                                           float *__restrict__ grad_x, float *__restrict__ grad_w,
                                                           const float *__restrict__ x_scale,
                                                           const float *__restrict__ x_shift, int x_act) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar[2];
    const int xs = k * c, wsz = k * kWD, ms = c * kWD, in_sz = xs + wsz + ms;
    float *dMt = smem + NBUF * in_sz;                         // [32][c], built from the arrived dM tile
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int g, int b) {
        float *base = smem + b * in_sz;
        mbar_expect_tx(&bar[b], (uint32_t)in_sz * 4u);
        bulk_g2s(base, x + (size_t)g * xs, (uint32_t)xs * 4u, &bar[b]);
        bulk_g2s(base + xs, w + (size_t)g * wsz, (uint32_t)wsz * 4u, &bar[b]);
        bulk_g2s(base + xs + wsz, grad_out + (size_t)g * ms, (uint32_t)ms * 4u, &bar[b]);
    };
    if (threadIdx.x == 0 && (int)blockIdx.x < groups) issue(blockIdx.x, 0);
    int it = 0;
    for (int g = blockIdx.x; g < groups; g += gridDim.x, ++it) {
        const int b = NBUF == 2 ? (it & 1) : 0, gn = g + gridDim.x;
        if (NBUF == 2 && threadIdx.x == 0 && gn < groups) issue(gn, b ^ 1);
        mbar_wait(&bar[b], (uint32_t)(NBUF == 2 ? ((it >> 1) & 1) : (it & 1)));
        const float *X = smem + b * in_sz, *W = X + xs, *dM = W + wsz;
        if (x_scale != nullptr) {
            bn_act_tile(smem + b * in_sz, xs, c, x_scale, x_shift, x_act);
            fence_async_smem();
        }
        for (int v = threadIdx.x; v < ms / 4; v += blockDim.x) {
            const float4 d = reinterpret_cast<const float4 *>(dM)[v];
            const int ch = v >> 3, w0 = (v & 7) * 4;           // element (ch, w0..w0+3)
            dMt[(w0 + 0) * c + ch] = d.x;
            dMt[(w0 + 1) * c + ch] = d.y;
            dMt[(w0 + 2) * c + ch] = d.z;
            dMt[(w0 + 3) * c + ch] = d.w;
        }
        __syncthreads();
        float *gx = grad_x + (size_t)g * xs;
        float *gw = grad_w + (size_t)g * wsz;
        const int ncg = c / 4, nxt = ncg * (k / 4), nwt = 8 * (k / 4);
        for (int tile = threadIdx.x; tile < nxt + nwt; tile += blockDim.x) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            if (tile < nxt) {
                // ---- dX[t][ch] = sum_w W[t][w] * dM[ch][w] : lanes walk channel groups (coalesced float4 stores)
                const int tg = tile / ncg, cg = tile - tg * ncg;
#pragma unroll 2
                for (int w0 = 0; w0 < kWD; w0 += 4) {
                    float wr[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 q = *reinterpret_cast<const float4 *>(W + (tg * 4 + i) * kWD + w0);
                        wr[i][0] = q.x; wr[i][1] = q.y; wr[i][2] = q.z; wr[i][3] = q.w;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 d = *reinterpret_cast<const float4 *>(dMt + (w0 + u) * c + cg * 4);
                        const float ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wr[i][u], ds[j], acc[i][j]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4 *>(gx + (tg * 4 + i) * c + cg * 4) =
                        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            } else {
                // ---- dW[t][w] = sum_ch X[t][ch] * dM[ch][w] : tiles of 4 t x 4 w
                const int wt = tile - nxt, wg = wt & 7, tg = wt >> 3;
#pragma unroll 2
                for (int c0 = 0; c0 < c; c0 += 4) {
                    float xr[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 q = *reinterpret_cast<const float4 *>(X + (tg * 4 + i) * c + c0);
                        xr[i][0] = q.x; xr[i][1] = q.y; xr[i][2] = q.z; xr[i][3] = q.w;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 d = *reinterpret_cast<const float4 *>(dM + (c0 + u) * kWD + wg * 4);
                        const float ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i][u], ds[j], acc[i][j]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4 *>(gw + (tg * 4 + i) * kWD + wg * 4) =
                        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            }
        }
        __syncthreads();
        if (NBUF == 1 && threadIdx.x == 0 && gn < groups) issue(gn, 0);   // single buffer: reload after everyone is done
    }
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

extern "C" int pasnl_pointconv_fwd_bn(int groups, int k, int c, const float *x, const float *w, float *out,
                                      const float *x_scale, const float *x_shift, int x_act, void *stream) {
    PASNL_CHECK_ARG((x_scale == nullptr) == (x_shift == nullptr) && x_act >= 0 && x_act <= 2 && aligned16(x_scale) &&
                    aligned16(x_shift));
    PASNL_CHECK_ARG(groups >= 0 && k >= 4 && k % 4 == 0 && k <= 256 && c >= 4 && c % 4 == 0 && c <= 1024);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(x && w && out && aligned16(x) && aligned16(w));
    const size_t smem = sizeof(float) * 2 * (size_t)(k * c + k * kWD);     // double-buffered X and W tiles
    PASNL_CHECK_ARG(smem <= 200 * 1024);
    cudaError_t e = cudaFuncSetAttribute(pointconv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    // one thread per 4x4 output tile, at most 256: C = 64 fills 128 threads
    int threads = 8 * (c / 4);
    threads = threads > kT ? kT : (threads + 31) / 32 * 32;
    const int grid = groups < 148 * 8 ? groups : 148 * 8;
    pointconv_fwd_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(groups, k, c, x, w, out, x_scale, x_shift, x_act);
    return pasnl_launch_status();
}

extern "C" int pasnl_pointconv_fwd(int groups, int k, int c, const float *x, const float *w, float *out, void *stream) {
    return pasnl_pointconv_fwd_bn(groups, k, c, x, w, out, nullptr, nullptr, 0, stream);
}

extern "C" int pasnl_pointconv_bwd_bn(int groups, int k, int c, const float *x, const float *w, const float *grad_out,
                                      float *grad_x, float *grad_w, const float *x_scale, const float *x_shift, int x_act,
                                      void *stream) {
    PASNL_CHECK_ARG((x_scale == nullptr) == (x_shift == nullptr) && x_act >= 0 && x_act <= 2 && aligned16(x_scale) &&
                    aligned16(x_shift));
    PASNL_CHECK_ARG(groups >= 0 && k >= 4 && k % 4 == 0 && k <= 256 && c >= 4 && c % 4 == 0 && c <= 1024);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(x && w && grad_out && grad_x && grad_w && aligned16(x) && aligned16(w) && aligned16(grad_out));
    const size_t in_bytes = sizeof(float) * (size_t)(k * c + k * kWD + c * kWD), t_bytes = sizeof(float) * (size_t)c * kWD;
    PASNL_CHECK_ARG(in_bytes + t_bytes <= 200 * 1024);
    // double-buffer the three input tiles while at least three CTAs still fit on an SM
    const bool dbl = 2 * in_bytes + t_bytes <= 72 * 1024;
    const size_t smem = (dbl ? 2 : 1) * in_bytes + t_bytes;
    const int grid = groups < 148 * 8 ? groups : 148 * 8;
    cudaError_t e;
    if (dbl) {
        e = cudaFuncSetAttribute(pointconv_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        pointconv_bwd_kernel<2><<<grid, kT, smem, (cudaStream_t)stream>>>(groups, k, c, x, w, grad_out, grad_x, grad_w, x_scale, x_shift,
                                                                         x_act);
    } else {
        e = cudaFuncSetAttribute(pointconv_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        pointconv_bwd_kernel<1><<<grid, kT, smem, (cudaStream_t)stream>>>(groups, k, c, x, w, grad_out, grad_x, grad_w, x_scale, x_shift,
                                                                         x_act);
    }
    return pasnl_launch_status();
}

extern "C" int pasnl_pointconv_bwd(int groups, int k, int c, const float *x, const float *w, const float *grad_out,
                                   float *grad_x, float *grad_w, void *stream) {
    return pasnl_pointconv_bwd_bn(groups, k, c, x, w, grad_out, grad_x, grad_w, nullptr, nullptr, 0, stream);
}
