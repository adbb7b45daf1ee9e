// adaptive_sampling.cu -- the point-set math of the Adaptive Sampling module (sm_100a).
//
// Replaces the chain of small TF ops in SampleWeights / AdaptiveSampling (reference utils/pointasnl_util.py:112-173):
// slice, batched k x k tf.matmul, softmax, batched matmul (:137-146) and softmax-over-k, tile, multiply, reduce_sum
// (:155,168-171).  In TF1 each is its own kernel round-tripping (B,S,k,.) tensors through HBM.
//   * as_core:     one WARP per neighbourhood; Q / Kf / V of the group are staged once in shared memory, the k x k
//                  logits, their softmax and W.V never leave the SM.  k <= 12 in every shipped config, so the group
//                  GEMMs are (12x65)x(65x12): far below a tcgen05 tile (M=64/128) -- CUDA-core fp32 by design.
//   * as_resample: one thread per (group, channel): softmax over the k members and the weighted sum fused.
// Both are HBM/launch bound: as_core moves 16*G*k*d bytes (+4*G*k*k for the saved W).
#include "common.cuh"

using namespace pasnl;

namespace {

// ---------------------------------------------------------------- SampleWeights attention core
// shared-memory layout per warp (floats): Q[k][dp] Kf[k][dp] V[k][dp] S[k][kp]   (fwd)
//                                         + dO[k][dp] dS[k][kp]                   (bwd)
// dp = d|1 (odd pitch: column walks of different rows hit different banks), kp = k+1.
__device__ __forceinline__ void warp_load_rows(float *dst, int dp, const float *src, int rows, int cols, int pitch,
                                               int lane) {
    for (int e = lane; e < rows * cols; e += 32) {
        const int r = e / cols, c = e - r * cols;
        dst[r * dp + c] = __ldg(src + (size_t)r * pitch + c);
    }
}

__global__ void as_core_fwd_kernel(int groups, int k, int d, const float *__restrict__ q, const float *__restrict__ kv,
                                   float scale, float *__restrict__ out, float *__restrict__ attn) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int dp = d | 1, kp = k + 1;
    const int per_warp = 3 * k * dp + k * kp;
    float *Q = smem + (size_t)warp * per_warp, *Kf = Q + k * dp, *V = Kf + k * dp, *S = V + k * dp;
    for (int g = blockIdx.x * nwarps + warp; g < groups; g += gridDim.x * nwarps) {
        warp_load_rows(Q, dp, q + (size_t)g * k * d, k, d, d, lane);
        warp_load_rows(Kf, dp, kv + (size_t)g * k * 2 * d, k, d, 2 * d, lane);
        warp_load_rows(V, dp, kv + (size_t)g * k * 2 * d + d, k, d, 2 * d, lane);
        __syncwarp();
        // logits: one (i,j) pair per lane per trip
        for (int e = lane; e < k * k; e += 32) {
            const int i = e / k, j = e - i * k;
            float acc = 0.f;
            for (int c = 0; c < d; ++c) acc = fmaf(Q[i * dp + c], Kf[j * dp + c], acc);
            S[i * kp + j] = acc * scale;
        }
        __syncwarp();
        // row softmax: lane i owns row i
        if (lane < k) {
            float mx = -INFINITY;
            for (int j = 0; j < k; ++j) mx = fmaxf(mx, S[lane * kp + j]);
            float sum = 0.f;
            for (int j = 0; j < k; ++j) {
                const float e = expf(S[lane * kp + j] - mx);
                S[lane * kp + j] = e;
                sum += e;
            }
            const float inv = 1.f / sum;
            for (int j = 0; j < k; ++j) {
                const float w = S[lane * kp + j] * inv;
                S[lane * kp + j] = w;
                attn[((size_t)g * k + lane) * k + j] = w;
            }
        }
        __syncwarp();
        // out = W V : lanes across the d columns
        for (int i = 0; i < k; ++i)
            for (int c = lane; c < d; c += 32) {
                float acc = 0.f;
                for (int j = 0; j < k; ++j) acc = fmaf(S[i * kp + j], V[j * dp + c], acc);
                out[((size_t)g * k + i) * d + c] = acc;
            }
        __syncwarp();
    }
}

__global__ void as_core_bwd_kernel(int groups, int k, int d, const float *__restrict__ q, const float *__restrict__ kv,
                                   const float *__restrict__ attn, float scale, const float *__restrict__ grad_out,
                                   float *__restrict__ grad_q, float *__restrict__ grad_kv) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int dp = d | 1, kp = k + 1;
    const int per_warp = 4 * k * dp + 2 * k * kp;
    float *Q = smem + (size_t)warp * per_warp, *Kf = Q + k * dp, *V = Kf + k * dp, *dO = V + k * dp;
    float *W = dO + k * dp, *dS = W + k * kp;
    for (int g = blockIdx.x * nwarps + warp; g < groups; g += gridDim.x * nwarps) {
        warp_load_rows(Q, dp, q + (size_t)g * k * d, k, d, d, lane);
        warp_load_rows(Kf, dp, kv + (size_t)g * k * 2 * d, k, d, 2 * d, lane);
        warp_load_rows(V, dp, kv + (size_t)g * k * 2 * d + d, k, d, 2 * d, lane);
        warp_load_rows(dO, dp, grad_out + (size_t)g * k * d, k, d, d, lane);
        warp_load_rows(W, kp, attn + (size_t)g * k * k, k, k, k, lane);
        __syncwarp();
        // dW[i][j] = dO_i . V_j
        for (int e = lane; e < k * k; e += 32) {
            const int i = e / k, j = e - i * k;
            float acc = 0.f;
            for (int c = 0; c < d; ++c) acc = fmaf(dO[i * dp + c], V[j * dp + c], acc);
            dS[i * kp + j] = acc;
        }
        __syncwarp();
        // softmax backward per row, folded with the logit scale: dS = scale * W * (dW - sum_j dW*W)
        if (lane < k) {
            float dot = 0.f;
            for (int j = 0; j < k; ++j) dot = fmaf(dS[lane * kp + j], W[lane * kp + j], dot);
            for (int j = 0; j < k; ++j) dS[lane * kp + j] = scale * W[lane * kp + j] * (dS[lane * kp + j] - dot);
        }
        __syncwarp();
        float *gq = grad_q + (size_t)g * k * d;
        float *gkv = grad_kv + (size_t)g * k * 2 * d;
        for (int r = 0; r < k; ++r)
            for (int c = lane; c < d; c += 32) {
                float aq = 0.f, ak = 0.f, av = 0.f;
                for (int j = 0; j < k; ++j) {
                    aq = fmaf(dS[r * kp + j], Kf[j * dp + c], aq);  // dQ_r  = sum_j dS[r][j] Kf_j
                    ak = fmaf(dS[j * kp + r], Q[j * dp + c], ak);   // dKf_r = sum_i dS[i][r] Q_i
                    av = fmaf(W[j * kp + r], dO[j * dp + c], av);   // dV_r  = sum_i W[i][r] dO_i
                }
                gq[(size_t)r * d + c] = aq;
                gkv[(size_t)r * 2 * d + c] = ak;
                gkv[(size_t)r * 2 * d + d + c] = av;
            }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- re-sampling: softmax over k + weighted sum
__global__ void __launch_bounds__(256) as_resample_fwd_kernel(int groups, int k, int cg, const float *__restrict__ logits,
                                                             const float *__restrict__ xyz, long xgs, long xms,
                                                             const float *__restrict__ feat, long fgs, long fms,
                                                             float *__restrict__ new_xyz, float *__restrict__ new_feat,
                                                             float *__restrict__ w) {
    const int cw = 1 + cg;
    const long total = (long)groups * cw;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long g = e / cw;
        const int ch = (int)(e - g * cw);
        const float *lg = logits + (size_t)g * k * cw + ch;
        float mx = -INFINITY;
        for (int t = 0; t < k; ++t) mx = fmaxf(mx, __ldg(lg + (size_t)t * cw));
        float sum = 0.f;
        for (int t = 0; t < k; ++t) sum += expf(__ldg(lg + (size_t)t * cw) - mx);
        const float inv = 1.f / sum;
        float *wo = w + (size_t)g * k * cw + ch;
        if (ch == 0) {
            float ax = 0.f, ay = 0.f, az = 0.f;
            for (int t = 0; t < k; ++t) {
                const float wt = expf(__ldg(lg + (size_t)t * cw) - mx) * inv;
                wo[(size_t)t * cw] = wt;
                const float *p = xyz + g * xgs + t * xms;
                ax = fmaf(wt, __ldg(p), ax);
                ay = fmaf(wt, __ldg(p + 1), ay);
                az = fmaf(wt, __ldg(p + 2), az);
            }
            new_xyz[g * 3 + 0] = ax;
            new_xyz[g * 3 + 1] = ay;
            new_xyz[g * 3 + 2] = az;
        } else {
            float acc = 0.f;
            for (int t = 0; t < k; ++t) {
                const float wt = expf(__ldg(lg + (size_t)t * cw) - mx) * inv;
                wo[(size_t)t * cw] = wt;
                acc = fmaf(wt, __ldg(feat + g * fgs + t * fms + (ch - 1)), acc);
            }
            new_feat[g * cg + (ch - 1)] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) as_resample_bwd_kernel(int groups, int k, int cg, const float *__restrict__ w,
                                                             const float *__restrict__ xyz, long xgs, long xms,
                                                             const float *__restrict__ feat, long fgs, long fms,
                                                             const float *__restrict__ g_xyz,
                                                             const float *__restrict__ g_feat,
                                                             float *__restrict__ grad_logits,
                                                             float *__restrict__ grad_xyz,
                                                             float *__restrict__ grad_feat) {
    const int cw = 1 + cg;
    const long total = (long)groups * cw;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long g = e / cw;
        const int ch = (int)(e - g * cw);
        const float *wi = w + (size_t)g * k * cw + ch;
        float *gl = grad_logits + (size_t)g * k * cw + ch;
        if (ch == 0) {
            const float gx = __ldg(g_xyz + g * 3), gy = __ldg(g_xyz + g * 3 + 1), gz = __ldg(g_xyz + g * 3 + 2);
            float dot = 0.f;
            for (int t = 0; t < k; ++t) {
                const float *p = xyz + g * xgs + t * xms;
                const float dw = fmaf(gz, __ldg(p + 2), fmaf(gy, __ldg(p + 1), gx * __ldg(p)));
                dot = fmaf(__ldg(wi + (size_t)t * cw), dw, dot);
            }
            for (int t = 0; t < k; ++t) {
                const float *p = xyz + g * xgs + t * xms;
                const float wt = __ldg(wi + (size_t)t * cw);
                const float dw = fmaf(gz, __ldg(p + 2), fmaf(gy, __ldg(p + 1), gx * __ldg(p)));
                gl[(size_t)t * cw] = wt * (dw - dot);
                float *o = grad_xyz + ((size_t)g * k + t) * 3;
                o[0] = gx * wt;
                o[1] = gy * wt;
                o[2] = gz * wt;
            }
        } else {
            const float gf = __ldg(g_feat + g * cg + (ch - 1));
            float dot = 0.f;
            for (int t = 0; t < k; ++t)
                dot = fmaf(__ldg(wi + (size_t)t * cw), gf * __ldg(feat + g * fgs + t * fms + (ch - 1)), dot);
            for (int t = 0; t < k; ++t) {
                const float wt = __ldg(wi + (size_t)t * cw);
                const float dw = gf * __ldg(feat + g * fgs + t * fms + (ch - 1));
                gl[(size_t)t * cw] = wt * (dw - dot);
                grad_feat[((size_t)g * k + t) * cg + (ch - 1)] = gf * wt;
            }
        }
    }
}

// warps per CTA that fit the shared-memory budget
int pick_warps(size_t per_warp_bytes) {
    const size_t budget = 160 * 1024;
    int w = (int)(budget / per_warp_bytes);
    if (w > 8) w = 8;
    return w;
}

}  // namespace

extern "C" int pasnl_as_core_fwd(int groups, int k, int d, const float *q, const float *kv, float scale, float *out,
                                 float *attn, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && k <= 32 && d >= 1 && d <= 256);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(q && kv && out && attn);
    const size_t per_warp = sizeof(float) * (size_t)(3 * k * (d | 1) + k * (k + 1));
    const int warps = pick_warps(per_warp);
    PASNL_CHECK_ARG(warps >= 1);
    const size_t smem = per_warp * warps;
    cudaError_t e = cudaFuncSetAttribute(as_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = ceil_div_i(groups, warps);
    if (grid > 148 * 16) grid = 148 * 16;
    as_core_fwd_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(groups, k, d, q, kv, scale, out, attn);
    return pasnl_launch_status();
}

extern "C" int pasnl_as_core_bwd(int groups, int k, int d, const float *q, const float *kv, const float *attn,
                                 float scale, const float *grad_out, float *grad_q, float *grad_kv, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && k <= 32 && d >= 1 && d <= 256);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(q && kv && attn && grad_out && grad_q && grad_kv);
    const size_t per_warp = sizeof(float) * (size_t)(4 * k * (d | 1) + 2 * k * (k + 1));
    const int warps = pick_warps(per_warp);
    PASNL_CHECK_ARG(warps >= 1);
    const size_t smem = per_warp * warps;
    cudaError_t e = cudaFuncSetAttribute(as_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int grid = ceil_div_i(groups, warps);
    if (grid > 148 * 16) grid = 148 * 16;
    as_core_bwd_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(groups, k, d, q, kv, attn, scale, grad_out,
                                                                         grad_q, grad_kv);
    return pasnl_launch_status();
}

extern "C" int pasnl_as_resample_fwd(int groups, int k, int cg, const float *logits, const float *xyz,
                                     long xyz_group_stride, long xyz_member_stride, const float *feat,
                                     long feat_group_stride, long feat_member_stride, float *new_xyz, float *new_feat,
                                     float *w, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && cg >= 0);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(logits && xyz && new_xyz && w && (cg == 0 || (feat && new_feat)));
    const long total = (long)groups * (1 + cg);
    int grid = ceil_div_i(total, 256);
    if (grid > 148 * 32) grid = 148 * 32;
    as_resample_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(groups, k, cg, logits, xyz, xyz_group_stride,
                                                                   xyz_member_stride, feat, feat_group_stride,
                                                                   feat_member_stride, new_xyz, new_feat, w);
    return pasnl_launch_status();
}

extern "C" int pasnl_as_resample_bwd(int groups, int k, int cg, const float *w, const float *xyz, long xyz_group_stride,
                                     long xyz_member_stride, const float *feat, long feat_group_stride,
                                     long feat_member_stride, const float *grad_new_xyz, const float *grad_new_feat,
                                     float *grad_logits, float *grad_xyz, float *grad_feat, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && cg >= 0);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(w && xyz && grad_new_xyz && grad_logits && grad_xyz &&
                    (cg == 0 || (feat && grad_new_feat && grad_feat)));
    const long total = (long)groups * (1 + cg);
    int grid = ceil_div_i(total, 256);
    if (grid > 148 * 32) grid = 148 * 32;
    as_resample_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        groups, k, cg, w, xyz, xyz_group_stride, xyz_member_stride, feat, feat_group_stride, feat_member_stride,
        grad_new_xyz, grad_new_feat, grad_logits, grad_xyz, grad_feat);
    return pasnl_launch_status();
}
