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

// Forward.  Thread tile = 4 channels x 4 weight columns: per neighbour t one LDS.128 of X[t][c..c+3] and one of
// W[t][w..w+3] feed 16 FMAs.  wg = t & 7 (8 groups of 4 weight columns), cg = t >> 3 (groups of 4 channels): a warp
// stores 4 full 128-byte rows of the (C x 32) output per instruction.
__global__ void __launch_bounds__(kT) pointconv_fwd_kernel(int groups, int k, int c, const float *__restrict__ x,
                                                           const float *__restrict__ w, float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float *X = smem, *W = X + k * c;
    const int wg = threadIdx.x & 7, cg0 = threadIdx.x >> 3;
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
        __syncthreads();
        load_tile(X, x + (size_t)g * k * c, k * c);
        load_tile(W, w + (size_t)g * k * kWD, k * kWD);
        __syncthreads();
        float *o = out + (size_t)g * c * kWD;
        for (int cg = cg0; cg < c / 4; cg += kT / 8) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int t = 0; t < k; ++t) {
                const float4 xv = *reinterpret_cast<const float4 *>(X + t * c + cg * 4);
                const float4 wv = *reinterpret_cast<const float4 *>(W + t * kWD + wg * 4);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xs[i], ws[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4 *>(o + (cg * 4 + i) * kWD + wg * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

// Backward.  dM is staged twice: row-major [ch][32] (for dW) and transposed [w][C] (for dX), so both products read
// float4 operands.  dX tile = 4 t x 4 ch, dW tile = 4 t x 4 w.
__global__ void __launch_bounds__(kT) pointconv_bwd_kernel(int groups, int k, int c, const float *__restrict__ x,
                                                           const float *__restrict__ w,
                                                           const float *__restrict__ grad_out,
                                                           float *__restrict__ grad_x, float *__restrict__ grad_w) {
    extern __shared__ __align__(16) float smem[];
    float *X = smem, *W = X + k * c, *dM = W + k * kWD, *dMt = dM + c * kWD;   // dM [c][32], dMt [32][c]
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
        __syncthreads();
        load_tile(X, x + (size_t)g * k * c, k * c);
        load_tile(W, w + (size_t)g * k * kWD, k * kWD);
        const float *go = grad_out + (size_t)g * c * kWD;
        for (int v = threadIdx.x; v < c * kWD / 4; v += kT) {
            const float4 d = __ldg(reinterpret_cast<const float4 *>(go) + v);
            reinterpret_cast<float4 *>(dM)[v] = d;
            const int ch = v >> 3, w0 = (v & 7) * 4;           // element (ch, w0..w0+3)
            dMt[(w0 + 0) * c + ch] = d.x;
            dMt[(w0 + 1) * c + ch] = d.y;
            dMt[(w0 + 2) * c + ch] = d.z;
            dMt[(w0 + 3) * c + ch] = d.w;
        }
        __syncthreads();
        // ---- dX[t][ch] = sum_w W[t][w] * dM[ch][w] : lanes walk channel groups (coalesced float4 stores)
        float *gx = grad_x + (size_t)g * k * c;
        const int ncg = c / 4;
        for (int tile = threadIdx.x; tile < ncg * (k / 4); tile += kT) {
            const int tg = tile / ncg, cg = tile - tg * ncg;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
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
        }
        // ---- dW[t][w] = sum_ch X[t][ch] * dM[ch][w] : tiles of 4 t x 4 w
        float *gw = grad_w + (size_t)g * k * kWD;
        for (int tile = threadIdx.x; tile < 8 * (k / 4); tile += kT) {
            const int wg = tile & 7, tg = tile >> 3;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
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
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

extern "C" int pasnl_pointconv_fwd(int groups, int k, int c, const float *x, const float *w, float *out, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 4 && k % 4 == 0 && k <= 256 && c >= 4 && c % 4 == 0 && c <= 1024);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(x && w && out && aligned16(x) && aligned16(w));
    const size_t smem = sizeof(float) * (size_t)(k * c + k * kWD);
    PASNL_CHECK_ARG(smem <= 200 * 1024);
    cudaError_t e = cudaFuncSetAttribute(pointconv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = groups < 148 * 8 ? groups : 148 * 8;
    pointconv_fwd_kernel<<<grid, kT, smem, (cudaStream_t)stream>>>(groups, k, c, x, w, out);
    return pasnl_launch_status();
}

extern "C" int pasnl_pointconv_bwd(int groups, int k, int c, const float *x, const float *w, const float *grad_out,
                                   float *grad_x, float *grad_w, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 4 && k % 4 == 0 && k <= 256 && c >= 4 && c % 4 == 0 && c <= 1024);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(x && w && grad_out && grad_x && grad_w && aligned16(x) && aligned16(w));
    const size_t smem = sizeof(float) * (size_t)(k * c + k * kWD + 2 * c * kWD);
    PASNL_CHECK_ARG(smem <= 200 * 1024);
    cudaError_t e = cudaFuncSetAttribute(pointconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = groups < 148 * 8 ? groups : 148 * 8;
    pointconv_bwd_kernel<<<grid, kT, smem, (cudaStream_t)stream>>>(groups, k, c, x, w, grad_out, grad_x, grad_w);
    return pasnl_launch_status();
}
