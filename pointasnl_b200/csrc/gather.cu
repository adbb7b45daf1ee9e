// gather.cu -- irregular row gathers and their scatter-add backwards (HBM-bound kernels).
//
// Replaces gatherpointKernel / scatteraddpointKernel (reference tf_ops/sampling/tf_sampling_g.cu:172-192),
// group_point_gpu / group_point_grad_gpu (tf_ops/grouping/tf_grouping_g.cu:40-78: one THREAD copies a whole
// c-float row scalar by scalar, <<<b,256>>>), the tf.gather_nd calls of sampling()/grouping()
// (utils/pointasnl_util.py:43-49,63-74) and threeinterpolate_cpu / _grad_cpu (tf_interpolate.cpp:107-153, one CPU
// thread).  Here: one thread per 16-byte (or 4-byte) piece of an OUTPUT row, so stores are fully coalesced and the
// source rows are read with the widest aligned vector the row pitch allows; grids cover all b*m*nsample rows.
// Algorithmic traffic per launch: 4*rows (idx) + 4*rows*c (out) + 4*b*n*c (source, read once; re-reads hit L2).
#include "common.cuh"

using namespace pasnl;

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// ---------------------------------------------------------------- generic row gather
// out[row, :] = src[(row / rows_per_batch) * n + idx[row], :],  row in [0, rows)
template <int VEC>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(long rows, int rows_per_batch, int n, int c,
                                                               const float *__restrict__ src,
                                                               const int *__restrict__ idx, float *__restrict__ out) {
    const int cv = c / VEC;
    const long total = rows * cv;
    for (long g = (long)blockIdx.x * kThreads + threadIdx.x; g < total; g += (long)gridDim.x * kThreads) {
        const long row = g / cv;
        const int col = (int)(g - row * cv) * VEC;
        const long bi = row / rows_per_batch;
        const long srow = bi * n + __ldg(idx + row);
        if constexpr (VEC == 4) {
            *reinterpret_cast<float4 *>(out + row * c + col) = ldg4(src + srow * c + col);
        } else {
            out[row * c + col] = __ldg(src + srow * c + col);
        }
    }
}

// grad_src[(row / rows_per_batch) * n + idx[row], :] += grad_out[row, :]
template <int VEC>
__global__ void __launch_bounds__(kThreads) scatter_rows_add_kernel(long rows, int rows_per_batch, int n, int c,
                                                                    const float *__restrict__ grad_out,
                                                                    const int *__restrict__ idx,
                                                                    float *__restrict__ grad_src) {
    const int cv = c / VEC;
    const long total = rows * cv;
    for (long g = (long)blockIdx.x * kThreads + threadIdx.x; g < total; g += (long)gridDim.x * kThreads) {
        const long row = g / cv;
        const int col = (int)(g - row * cv) * VEC;
        const long bi = row / rows_per_batch;
        const long srow = bi * n + __ldg(idx + row);
        if constexpr (VEC == 4) {
            const float4 v = ldg4(grad_out + row * c + col);
            atomicAdd(reinterpret_cast<float4 *>(grad_src + srow * c + col), v);  // red.global.add.v4.f32 (sm_90+)
        } else {
            atomicAdd(grad_src + srow * c + col, __ldg(grad_out + row * c + col));
        }
    }
}

int grid_for(long work_items) {
    long blocks = (work_items + kThreads - 1) / kThreads;
    const long cap = 148L * 32;  // grid-stride beyond ~32 CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

int gather_rows(long rows, int rows_per_batch, int n, int c, const float *src, const int *idx, float *out,
                cudaStream_t st) {
    if (rows == 0 || c == 0) return PASNL_OK;
    if (c % 4 == 0 && aligned16(src) && aligned16(out)) {
        gather_rows_kernel<4><<<grid_for(rows * (c / 4)), kThreads, 0, st>>>(rows, rows_per_batch, n, c, src, idx, out);
    } else {
        gather_rows_kernel<1><<<grid_for(rows * c), kThreads, 0, st>>>(rows, rows_per_batch, n, c, src, idx, out);
    }
    return pasnl_launch_status();
}

int scatter_rows_add(long rows, int rows_per_batch, int n, int c, long src_rows, const float *grad_out, const int *idx,
                     float *grad_src, cudaStream_t st) {
    if (src_rows * c > 0) {
        cudaError_t e = cudaMemsetAsync(grad_src, 0, sizeof(float) * (size_t)src_rows * c, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (rows == 0 || c == 0) return PASNL_OK;
    if (c % 4 == 0 && aligned16(grad_out) && aligned16(grad_src)) {
        scatter_rows_add_kernel<4>
            <<<grid_for(rows * (c / 4)), kThreads, 0, st>>>(rows, rows_per_batch, n, c, grad_out, idx, grad_src);
    } else {
        scatter_rows_add_kernel<1>
            <<<grid_for(rows * c), kThreads, 0, st>>>(rows, rows_per_batch, n, c, grad_out, idx, grad_src);
    }
    return pasnl_launch_status();
}

// ---------------------------------------------------------------- fused SA-layer neighbourhood gather
// out[row, :] = [ xyz[i] - center[q] | xyz[i] | feat[i] ]   (leading part only if center != nullptr)
// One WARP per output row: the row is decoded once (32-bit arithmetic), lanes then stride over its columns, so
// both the feature read (one contiguous source row) and the output write (one contiguous output row) are coalesced.
// (r01 used one thread per output float with 64-bit div/mod each: 130 instructions per float, 13 % of HBM peak.)
// LPR = lanes per row (a power of two <= 32): narrow rows (cls layer 1: 9 floats) share a warp, wide rows get all 32 lanes.
template <int LPR>
__global__ void __launch_bounds__(kThreads) group_concat_kernel(int rows, int m, int nsample, int n, int c,
                                                                int idx_stride, const float *__restrict__ xyz,
                                                                const float *__restrict__ feat,
                                                                const int *__restrict__ idx,
                                                                const float *__restrict__ center,
                                                                float *__restrict__ out) {
    constexpr int RPW = 32 / LPR;   // rows per warp per trip
    const int lead = center ? 3 : 0;
    const int w = lead + 3 + c;
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
    for (int row = warp * RPW + grp; row < rows; row += nwarps * RPW) {
        const int q = row / nsample, t = row - q * nsample;
        const int bi = q / m;
        const size_t srow = (size_t)bi * n + __ldg(idx + (size_t)q * idx_stride + t);
        float *o = out + (size_t)row * w;
        const float *f = feat + srow * c;
        for (int col = sub; col < w; col += LPR) {
            float v;
            if (col < lead) v = __fsub_rn(__ldg(xyz + srow * 3 + col), __ldg(center + (size_t)q * 3 + col));
            else if (col < lead + 3) v = __ldg(xyz + srow * 3 + (col - lead));
            else v = __ldg(f + (col - lead - 3));
            o[col] = v;
        }
    }
}

// Backward: LPR lanes per grad_out row.  Relative and absolute xyz parts both land in grad_xyz[i];
// grad_center[q] accumulates minus the relative part; the feature part is a coalesced row of reductions.
template <int LPR, bool VEC4>
__global__ void __launch_bounds__(kThreads) group_concat_grad_kernel(int rows, int m, int nsample, int n, int c,
                                                                     int idx_stride, int has_center,
                                                                     const float *__restrict__ grad_out,
                                                                     const int *__restrict__ idx,
                                                                     float *__restrict__ grad_xyz,
                                                                     float *__restrict__ grad_feat,
                                                                     float *__restrict__ grad_center) {
    constexpr int RPW = 32 / LPR;
    const int lead = has_center ? 3 : 0;
    const int w = lead + 3 + c;
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
    for (int row = warp * RPW + grp; row < rows; row += nwarps * RPW) {
        const int q = row / nsample, t = row - q * nsample;
        const int bi = q / m;
        const size_t srow = (size_t)bi * n + __ldg(idx + (size_t)q * idx_stride + t);
        const float *g = grad_out + (size_t)row * w;
        for (int col = sub; col < 3; col += LPR) {
            float v = __ldg(g + lead + col);
            if (lead) {
                const float r = __ldg(g + col);
                atomicAdd(grad_center + (size_t)q * 3 + col, -r);
                v += r;
            }
            atomicAdd(grad_xyz + srow * 3 + col, v);
        }
        const float *gf = g + lead + 3;
        float *dst = grad_feat + srow * c;
        if constexpr (VEC4) {
            // destination rows are 16-byte aligned (c % 4 == 0); the source row is only 4/8-byte aligned
            for (int v4 = sub; v4 < c / 4; v4 += LPR) {
                float4 v;
                v.x = __ldg(gf + v4 * 4 + 0);
                v.y = __ldg(gf + v4 * 4 + 1);
                v.z = __ldg(gf + v4 * 4 + 2);
                v.w = __ldg(gf + v4 * 4 + 3);
                atomicAdd(reinterpret_cast<float4 *>(dst + v4 * 4), v);   // red.global.add.v4.f32
            }
        } else {
            for (int col = sub; col < c; col += LPR) atomicAdd(dst + col, __ldg(gf + col));
        }
    }
}

int warp_grid(long rows, int rows_per_warp) {
    long warps = (rows + rows_per_warp - 1) / rows_per_warp;
    long blocks = (warps + (kThreads / 32) - 1) / (kThreads / 32);
    const long cap = 148L * 16;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// ---------------------------------------------------------------- max over the neighbour axis
// out[g, c] = max_k x[g, k, c]  (tf.reduce_max(new_point, axis=[2]), utils/pointasnl_util.py:258) and its gradient, which
// TF splits evenly among tied maxima (indicators / num_selected).  One thread per (group, channel): consecutive threads
// read consecutive channels, so every k-row is a coalesced access.  HBM bound: 4*G*K*C bytes forward, 8*G*K*C backward.
__global__ void __launch_bounds__(kThreads) maxpool_k_kernel(long gc, int k, int c, const float *__restrict__ x,
                                                             float *__restrict__ out) {
    for (long e = (long)blockIdx.x * kThreads + threadIdx.x; e < gc; e += (long)gridDim.x * kThreads) {
        const long g = e / c;
        const int ch = (int)(e - g * c);
        const float *p = x + g * (long)k * c + ch;
        float m = __ldg(p);
        for (int t = 1; t < k; ++t) m = fmaxf(m, __ldg(p + (long)t * c));
        out[e] = m;
    }
}
__global__ void __launch_bounds__(kThreads) maxpool_k_grad_kernel(long gc, int k, int c, const float *__restrict__ x,
                                                                  const float *__restrict__ out,
                                                                  const float *__restrict__ grad_out,
                                                                  float *__restrict__ grad_x) {
    for (long e = (long)blockIdx.x * kThreads + threadIdx.x; e < gc; e += (long)gridDim.x * kThreads) {
        const long g = e / c;
        const int ch = (int)(e - g * c);
        const float *p = x + g * (long)k * c + ch;
        float *q = grad_x + g * (long)k * c + ch;
        const float m = __ldg(out + e);
        int cnt = 0;
        for (int t = 0; t < k; ++t) cnt += (__ldg(p + (long)t * c) == m) ? 1 : 0;
        const float share = __ldg(grad_out + e) / (float)cnt;
        for (int t = 0; t < k; ++t) q[(long)t * c] = (__ldg(p + (long)t * c) == m) ? share : 0.f;
    }
}

// Few groups with many neighbours (the global pooling of the classifier: 32 clouds x 512 points x 512 channels): one
// thread per output would leave 16 K threads each walking 512 strided rows.  Here a CTA = 32 channels x 8 k-slices:
// lanes read 128 contiguous bytes per row, the 8 slices are combined through shared memory.
constexpr int kPoolSlices = 8;
__global__ void __launch_bounds__(32 * kPoolSlices) maxpool_k_sliced_kernel(int k, int c, const float *__restrict__ x,
                                                                          float *__restrict__ out) {
    __shared__ float red[kPoolSlices][32];
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + lane;
    const long g = blockIdx.y;
    float m = -INFINITY;
    if (ch < c) {
        const float *p = x + g * (long)k * c + ch;
#pragma unroll 4
        for (int t = sl; t < k; t += kPoolSlices) m = fmaxf(m, __ldg(p + (long)t * c));
    }
    red[sl][lane] = m;
    __syncthreads();
    if (sl == 0 && ch < c) {
#pragma unroll
        for (int i = 1; i < kPoolSlices; ++i) m = fmaxf(m, red[i][lane]);
        out[g * c + ch] = m;
    }
}
__global__ void __launch_bounds__(32 * kPoolSlices) maxpool_k_grad_sliced_kernel(int k, int c, const float *__restrict__ x,
                                                                               const float *__restrict__ out,
                                                                               const float *__restrict__ grad_out,
                                                                               float *__restrict__ grad_x) {
    __shared__ int cnt_s[kPoolSlices][32];
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + lane;
    const long g = blockIdx.y;
    const bool ok = ch < c;
    const float *p = x + g * (long)k * c + ch;
    const float m = ok ? __ldg(out + g * c + ch) : 0.f;
    int cnt = 0;
    if (ok) {
#pragma unroll 4
        for (int t = sl; t < k; t += kPoolSlices) cnt += (__ldg(p + (long)t * c) == m) ? 1 : 0;
    }
    cnt_s[sl][lane] = cnt;
    __syncthreads();
    if (!ok) return;
    int total = 0;
#pragma unroll
    for (int i = 0; i < kPoolSlices; ++i) total += cnt_s[i][lane];
    const float share = __ldg(grad_out + g * c + ch) / (float)total;
    float *q = grad_x + g * (long)k * c + ch;
#pragma unroll 4
    for (int t = sl; t < k; t += kPoolSlices) q[(long)t * c] = (__ldg(p + (long)t * c) == m) ? share : 0.f;
}

// ---------------------------------------------------------------- three_interpolate
// out[row, :] = (p1*w1 + p2*w2) + p3*w3 with separate roundings (the x86 reference has no FMA): bit-exact.
template <int VEC>
__global__ void __launch_bounds__(kThreads) three_interp_kernel(long rows, int n, int m, int c,
                                                                const float *__restrict__ points,
                                                                const int *__restrict__ idx,
                                                                const float *__restrict__ weight,
                                                                float *__restrict__ out) {
    const int cv = c / VEC;
    const long total = rows * cv;
    for (long g = (long)blockIdx.x * kThreads + threadIdx.x; g < total; g += (long)gridDim.x * kThreads) {
        const long row = g / cv;
        const int col = (int)(g - row * cv) * VEC;
        const long bi = row / n;
        const float w1 = __ldg(weight + row * 3), w2 = __ldg(weight + row * 3 + 1), w3 = __ldg(weight + row * 3 + 2);
        const float *p1 = points + (bi * m + __ldg(idx + row * 3)) * c + col;
        const float *p2 = points + (bi * m + __ldg(idx + row * 3 + 1)) * c + col;
        const float *p3 = points + (bi * m + __ldg(idx + row * 3 + 2)) * c + col;
        if constexpr (VEC == 4) {
            const float4 a = ldg4(p1), b4 = ldg4(p2), d = ldg4(p3);
            float4 r;
            r.x = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w1), __fmul_rn(b4.x, w2)), __fmul_rn(d.x, w3));
            r.y = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w1), __fmul_rn(b4.y, w2)), __fmul_rn(d.y, w3));
            r.z = __fadd_rn(__fadd_rn(__fmul_rn(a.z, w1), __fmul_rn(b4.z, w2)), __fmul_rn(d.z, w3));
            r.w = __fadd_rn(__fadd_rn(__fmul_rn(a.w, w1), __fmul_rn(b4.w, w2)), __fmul_rn(d.w, w3));
            *reinterpret_cast<float4 *>(out + row * c + col) = r;
        } else {
            out[row * c + col] =
                __fadd_rn(__fadd_rn(__fmul_rn(__ldg(p1), w1), __fmul_rn(__ldg(p2), w2)), __fmul_rn(__ldg(p3), w3));
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) three_interp_grad_kernel(long rows, int n, int m, int c,
                                                                     const float *__restrict__ grad_out,
                                                                     const int *__restrict__ idx,
                                                                     const float *__restrict__ weight,
                                                                     float *__restrict__ grad_points) {
    const int cv = c / VEC;
    const long total = rows * cv;
    for (long g = (long)blockIdx.x * kThreads + threadIdx.x; g < total; g += (long)gridDim.x * kThreads) {
        const long row = g / cv;
        const int col = (int)(g - row * cv) * VEC;
        const long bi = row / n;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const float w = __ldg(weight + row * 3 + t);
            float *dst = grad_points + (bi * m + __ldg(idx + row * 3 + t)) * c + col;
            if constexpr (VEC == 4) {
                float4 v = ldg4(grad_out + row * c + col);
                v.x = __fmul_rn(v.x, w); v.y = __fmul_rn(v.y, w); v.z = __fmul_rn(v.z, w); v.w = __fmul_rn(v.w, w);
                atomicAdd(reinterpret_cast<float4 *>(dst), v);
            } else {
                atomicAdd(dst, __fmul_rn(__ldg(grad_out + row * c + col), w));
            }
        }
    }
}

}  // namespace

// ================================================================ C ABI
extern "C" int pasnl_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if ((long)b * m == 0) return PASNL_OK;
    PASNL_CHECK_ARG(inp && idx && out && n >= 1);
    return gather_rows((long)b * m, m, n, 3, inp, idx, out, (cudaStream_t)stream);
}

extern "C" int pasnl_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g,
                                       void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if ((long)b * n == 0) return PASNL_OK;
    PASNL_CHECK_ARG(inp_g && (m == 0 || (out_g && idx)));
    return scatter_rows_add((long)b * m, m > 0 ? m : 1, n, 3, (long)b * n, out_g, idx, inp_g, (cudaStream_t)stream);
}

extern "C" int pasnl_gather_rows(int b, int n, int c, int m, const float *src, const int *idx, float *out,
                                 void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0);
    if ((long)b * m * c == 0) return PASNL_OK;
    PASNL_CHECK_ARG(src && idx && out && n >= 1);
    return gather_rows((long)b * m, m, n, c, src, idx, out, (cudaStream_t)stream);
}

extern "C" int pasnl_gather_rows_grad(int b, int n, int c, int m, const float *out_g, const int *idx, float *src_g,
                                      void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0);
    if ((long)b * n * c == 0) return PASNL_OK;
    PASNL_CHECK_ARG(src_g && (m == 0 || (out_g && idx)));
    return scatter_rows_add((long)b * m, m > 0 ? m : 1, n, c, (long)b * n, out_g, idx, src_g, (cudaStream_t)stream);
}

extern "C" int pasnl_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                                 float *out, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0 && nsample >= 0);
    const long rows = (long)b * m * nsample;
    if (rows * c == 0) return PASNL_OK;
    PASNL_CHECK_ARG(points && idx && out && n >= 1);
    return gather_rows(rows, m * nsample, n, c, points, idx, out, (cudaStream_t)stream);
}

extern "C" int pasnl_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                                      float *grad_points, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0 && nsample >= 0);
    if ((long)b * n * c == 0) return PASNL_OK;
    const long rows = (long)b * m * nsample;
    PASNL_CHECK_ARG(grad_points && (rows == 0 || (grad_out && idx)));
    return scatter_rows_add(rows, rows > 0 ? m * nsample : 1, n, c, (long)b * n, grad_out, idx, grad_points,
                            (cudaStream_t)stream);
}

extern "C" int pasnl_group_concat(int b, int n, int c, int m, int nsample, const float *xyz, const float *feat,
                                  const int *idx, int idx_stride, const float *center, float *out, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0 && nsample >= 0 && idx_stride >= nsample);
    const long rows = (long)b * m * nsample;
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(xyz && idx && out && n >= 1 && (c == 0 || feat));
    PASNL_CHECK_ARG(rows < (1L << 31));
    const int w = (center ? 3 : 0) + 3 + c;
    cudaStream_t st = (cudaStream_t)stream;
#define GC_LAUNCH(LPR_) group_concat_kernel<LPR_><<<warp_grid(rows, 32 / LPR_), kThreads, 0, st>>>( \
        (int)rows, m, nsample, n, c, idx_stride, xyz, feat, idx, center, out)
    if (w <= 8) GC_LAUNCH(4);
    else if (w <= 16) GC_LAUNCH(8);
    else if (w <= 48) GC_LAUNCH(16);
    else GC_LAUNCH(32);
#undef GC_LAUNCH
    return pasnl_launch_status();
}

extern "C" int pasnl_group_concat_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                                       int idx_stride, int has_center, float *grad_xyz, float *grad_feat,
                                       float *grad_center, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0 && nsample >= 0 && idx_stride >= nsample);
    PASNL_CHECK_ARG(grad_xyz || (long)b * n == 0);
    PASNL_CHECK_ARG(!has_center || grad_center || (long)b * m == 0);
    PASNL_CHECK_ARG(c == 0 || grad_feat || (long)b * n == 0);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if ((long)b * n > 0) {
        e = cudaMemsetAsync(grad_xyz, 0, sizeof(float) * (size_t)b * n * 3, st);
        if (e != cudaSuccess) return (int)e;
        if (c > 0) {
            e = cudaMemsetAsync(grad_feat, 0, sizeof(float) * (size_t)b * n * c, st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (has_center && (long)b * m > 0) {
        e = cudaMemsetAsync(grad_center, 0, sizeof(float) * (size_t)b * m * 3, st);
        if (e != cudaSuccess) return (int)e;
    }
    const long rows = (long)b * m * nsample;
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(grad_out && idx && n >= 1);
    PASNL_CHECK_ARG(rows < (1L << 31));
    const int w = (has_center ? 3 : 0) + 3 + c;
#define GCG_LAUNCH(LPR_, V_) group_concat_grad_kernel<LPR_, V_><<<warp_grid(rows, 32 / LPR_), kThreads, 0, st>>>( \
        (int)rows, m, nsample, n, c, idx_stride, has_center, grad_out, idx, grad_xyz, grad_feat, grad_center)
    if (c > 0 && c % 4 == 0 && aligned16(grad_feat)) {
        if (w <= 48) GCG_LAUNCH(8, true);
        else GCG_LAUNCH(32, true);
    } else {
        if (w <= 8) GCG_LAUNCH(4, false);
        else if (w <= 16) GCG_LAUNCH(8, false);
        else if (w <= 48) GCG_LAUNCH(16, false);
        else GCG_LAUNCH(32, false);
    }
#undef GCG_LAUNCH
    return pasnl_launch_status();
}

extern "C" int pasnl_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                                       const float *weight, float *out, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0);
    const long rows = (long)b * n;
    if (rows * c == 0) return PASNL_OK;
    PASNL_CHECK_ARG(points && idx && weight && out && m >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
        three_interp_kernel<4><<<grid_for(rows * (c / 4)), kThreads, 0, st>>>(rows, n, m, c, points, idx, weight, out);
    } else {
        three_interp_kernel<1><<<grid_for(rows * c), kThreads, 0, st>>>(rows, n, m, c, points, idx, weight, out);
    }
    return pasnl_launch_status();
}

extern "C" int pasnl_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                            const float *weight, float *grad_points, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && c >= 0);
    if ((long)b * m * c == 0) return PASNL_OK;
    PASNL_CHECK_ARG(grad_points);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, st);
    if (e != cudaSuccess) return (int)e;
    const long rows = (long)b * n;
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(grad_out && idx && weight);
    if (c % 4 == 0 && aligned16(grad_out) && aligned16(grad_points)) {
        three_interp_grad_kernel<4>
            <<<grid_for(rows * (c / 4)), kThreads, 0, st>>>(rows, n, m, c, grad_out, idx, weight, grad_points);
    } else {
        three_interp_grad_kernel<1>
            <<<grid_for(rows * c), kThreads, 0, st>>>(rows, n, m, c, grad_out, idx, weight, grad_points);
    }
    return pasnl_launch_status();
}

extern "C" int pasnl_maxpool_k(long groups, int k, int c, const float *x, float *out, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && c >= 1);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(x && out);
    if (groups * c < (1L << 18) && k >= 8 * kPoolSlices && groups <= 65535) {
        dim3 grid((unsigned)((c + 31) / 32), (unsigned)groups);
        maxpool_k_sliced_kernel<<<grid, 32 * kPoolSlices, 0, (cudaStream_t)stream>>>(k, c, x, out);
        return pasnl_launch_status();
    }
    maxpool_k_kernel<<<grid_for(groups * c), kThreads, 0, (cudaStream_t)stream>>>(groups * c, k, c, x, out);
    return pasnl_launch_status();
}

extern "C" int pasnl_maxpool_k_grad(long groups, int k, int c, const float *x, const float *out, const float *grad_out,
                                    float *grad_x, void *stream) {
    PASNL_CHECK_ARG(groups >= 0 && k >= 1 && c >= 1);
    if (groups == 0) return PASNL_OK;
    PASNL_CHECK_ARG(x && out && grad_out && grad_x);
    if (groups * c < (1L << 18) && k >= 8 * kPoolSlices && groups <= 65535) {
        dim3 grid((unsigned)((c + 31) / 32), (unsigned)groups);
        maxpool_k_grad_sliced_kernel<<<grid, 32 * kPoolSlices, 0, (cudaStream_t)stream>>>(k, c, x, out, grad_out, grad_x);
        return pasnl_launch_status();
    }
    maxpool_k_grad_kernel<<<grid_for(groups * c), kThreads, 0, (cudaStream_t)stream>>>(groups * c, k, c, x, out,
                                                                                        grad_out, grad_x);
    return pasnl_launch_status();
}
