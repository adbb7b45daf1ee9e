// nonlocal.cu -- attention core of the Point Non-Local cell, flash style, fp32 (sm_100a).
//
// Replaces, in PointNonLocalCell mode='dot' (reference utils/pointasnl_util.py:195-212), the sequence
//   tf.matmul(Q, Kf^T) -> / sqrt(D) -> tf.nn.softmax -> tf.matmul(A, V)
// which materialises the (B, npoint, ndataset) attention map A in HBM and reads/writes it >= 3 times
// (67 MB at cls layer 1, 210 MB at ScanNet layer1_1, 4.3 GB at SemanticKITTI layer1_1).
// Here A never exists: a CTA owns a 64-row tile, streams the other side in tiles through shared memory, keeps the
// online-softmax state in registers (forward) or recomputes P from the saved row log-sum-exp (backward).
//
// Arithmetic is fp32 FMA on CUDA cores on purpose in this revision: the parity bar is 1e-4 relative against the
// fp32 reference and a single-pass TF32 tensor-core product (10-bit mantissa) gives ~1e-3 on the logits.  The
// tcgen05 path for the big shapes (KITTI layer1_1: 137 GFLOP) has to be a 3xTF32 / bf16x3 split and is tracked in
// DESIGN.md as the next step for this kernel; the tile/LSE structure below is what it will reuse.
//
// Thread layout: 256 threads = 16 (ty) x 16 (tx); thread (ty,tx) owns rows ty*4..+3 and, for the score tile,
// columns tx*RN..+RN-1.  Operand tiles for the score products are stored TRANSPOSED in shared memory ([d][row]) so
// that one LDS.128 fetches the 4 rows / 4 columns a thread needs for a given d.
#include <cstdlib>

#include "common.cuh"

using namespace pasnl;

namespace {

constexpr int BM = 64;
constexpr int NT = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// dst[c * ROWS + r] = src[r * pitch + c], zero-filled outside nvalid x d.
template <int D, int ROWS>
__device__ __forceinline__ void load_transposed(float *dst, const float *__restrict__ src, int nvalid, int d, int pitch) {
    constexpr int CG = NT / ROWS;
    const int r = threadIdx.x % ROWS;
    for (int c = (threadIdx.x / ROWS) * 4; c < D; c += CG * 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nvalid && c < d) v = ldg4(src + (size_t)r * pitch + c);
        dst[(c + 0) * ROWS + r] = v.x;
        dst[(c + 1) * ROWS + r] = v.y;
        dst[(c + 2) * ROWS + r] = v.z;
        dst[(c + 3) * ROWS + r] = v.w;
    }
}
// dst[r * D + c] = src[r * pitch + c], zero-filled outside nvalid x d.
template <int D, int ROWS>
__device__ __forceinline__ void load_rowmajor(float *dst, const float *__restrict__ src, int nvalid, int d, int pitch) {
    constexpr int VPR = D / 4;
    for (int v = threadIdx.x; v < ROWS * VPR; v += NT) {
        const int r = v / VPR, c = (v - r * VPR) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nvalid && c < d) x = ldg4(src + (size_t)r * pitch + c);
        *reinterpret_cast<float4 *>(dst + r * D + c) = x;
    }
}

// acc[i][j] += sum_dd At[dd][ty*4+i] * Bt[dd][tx*RN+j]
template <int D, int ROWS_B, int RN>
__device__ __forceinline__ void tile_dot(float (&acc)[4][RN], const float *At, const float *Bt, int ty, int tx) {
#pragma unroll 8
    for (int dd = 0; dd < D; ++dd) {
        const float4 a = *reinterpret_cast<const float4 *>(At + dd * BM + ty * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
        float bv[RN];
        if constexpr (RN == 4) {
            const float4 b = *reinterpret_cast<const float4 *>(Bt + dd * ROWS_B + tx * 4);
            bv[0] = b.x; bv[1] = b.y; bv[2] = b.z; bv[3] = b.w;
        } else {
            const float2 b = *reinterpret_cast<const float2 *>(Bt + dd * ROWS_B + tx * 2);
            bv[0] = b.x; bv[1] = b.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// column owned by (tx, c) in the [rows][D] output tiles: D/16 columns per thread
template <int D>
__device__ __forceinline__ int out_col(int tx, int c) {
    if constexpr (D / 16 >= 4) return (c >> 2) * 64 + tx * 4 + (c & 3);
    else return tx * (D / 16) + c;
}

// acc[i][c] += sum_nn Ps[ty*4+i][nn] * Vs[nn][out_col(tx,c)]   (Ps pitch BN+4, Vs pitch D)
template <int D, int BN>
__device__ __forceinline__ void tile_pv(float (&acc)[4][D / 16], const float *Ps, const float *Vs, int ty, int tx) {
    constexpr int CPT = D / 16, PP = BN + 4;
#pragma unroll 2
    for (int nn = 0; nn < BN; nn += 4) {
        float pv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 p = *reinterpret_cast<const float4 *>(Ps + (ty * 4 + i) * PP + nn);
            pv[i][0] = p.x; pv[i][1] = p.y; pv[i][2] = p.z; pv[i][3] = p.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float vv[CPT];
            if constexpr (CPT == 2) {
                const float2 v = *reinterpret_cast<const float2 *>(Vs + (nn + u) * D + tx * 2);
                vv[0] = v.x; vv[1] = v.y;
            } else {
#pragma unroll
                for (int gq = 0; gq < CPT / 4; ++gq) {
                    const float4 v = *reinterpret_cast<const float4 *>(Vs + (nn + u) * D + gq * 64 + tx * 4);
                    vv[gq * 4 + 0] = v.x; vv[gq * 4 + 1] = v.y; vv[gq * 4 + 2] = v.z; vv[gq * 4 + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(pv[i][u], vv[c], acc[i][c]);
        }
    }
}

__device__ __forceinline__ float half_warp_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, 1));
    v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, 2));
    v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, 4));
    v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, 8));
    return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
    v += __shfl_xor_sync(FULL_MASK, v, 1);
    v += __shfl_xor_sync(FULL_MASK, v, 2);
    v += __shfl_xor_sync(FULL_MASK, v, 4);
    v += __shfl_xor_sync(FULL_MASK, v, 8);
    return v;
}

// ---------------------------------------------------------------- forward
template <int D>
__global__ void __launch_bounds__(NT) nl_fwd_kernel(int s, int n, int d, const float *__restrict__ q,
                                                    const float *__restrict__ kv, float scale_log2,
                                                    float *__restrict__ out, float *__restrict__ lse) {
    constexpr int BN = 64, CPT = D / 16, PP = BN + 4;
    extern __shared__ __align__(16) float smem[];
    float *Qt = smem, *Kt = Qt + D * BM, *Vs = Kt + D * BN, *Ps = Vs + BN * D;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int bi = blockIdx.y, row0 = blockIdx.x * BM;
    const int rows_valid = min(BM, s - row0);
    load_transposed<D, BM>(Qt, q + ((size_t)bi * s + row0) * d, rows_valid, d, d);

    float m_i[4], l_i[4], acc[4][CPT];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_i[i] = -INFINITY;
        l_i[i] = 0.f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;
    }
    for (int col0 = 0; col0 < n; col0 += BN) {
        const int cv = min(BN, n - col0);
        __syncthreads();
        const float *kvt = kv + ((size_t)bi * n + col0) * 2 * d;
        load_transposed<D, BN>(Kt, kvt, cv, d, 2 * d);
        load_rowmajor<D, BN>(Vs, kvt + d, cv, d, 2 * d);
        __syncthreads();
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
        tile_dot<D, BN, 4>(sc, Qt, Kt, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sc[i][j] = (tx * 4 + j < cv) ? sc[i][j] * scale_log2 : -INFINITY;
                mx = fmaxf(mx, sc[i][j]);
            }
            mx = half_warp_max(mx);
            const float mnew = fmaxf(m_i[i], mx);
            const float alpha = exp2f(m_i[i] - mnew);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sc[i][j] = exp2f(sc[i][j] - mnew);
                rs += sc[i][j];
            }
            rs = half_warp_sum(rs);
            l_i[i] = fmaf(l_i[i], alpha, rs);
            m_i[i] = mnew;
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] *= alpha;
            *reinterpret_cast<float4 *>(Ps + (ty * 4 + i) * PP + tx * 4) =
                make_float4(sc[i][0], sc[i][1], sc[i][2], sc[i][3]);
        }
        __syncthreads();
        tile_pv<D, BN>(acc, Ps, Vs, ty, tx);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + ty * 4 + i;
        if (r < s) {
            const float inv = 1.f / l_i[i];
            float *o = out + ((size_t)bi * s + r) * d;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = out_col<D>(tx, c);
                if (col < d) o[col] = acc[i][c] * inv;
            }
            if (tx == 0) lse[(size_t)bi * s + r] = (m_i[i] + log2f(l_i[i])) * kLn2;
        }
    }
}

// ---------------------------------------------------------------- backward
// delta[row] = sum_c dO[row][c] * O[row][c]
__global__ void nl_delta_kernel(long rows, int d, const float *__restrict__ dout, const float *__restrict__ out,
                                float *__restrict__ delta) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float acc = 0.f;
    for (int c = 0; c < d; c += 4) {
        const float4 a = ldg4(dout + r * d + c), b = ldg4(out + r * d + c);
        acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
    }
    delta[r] = acc;
}

// KEYS_OUTER = false: CTA owns 64 queries, streams key tiles, produces dQ = dS K.
// KEYS_OUTER = true : CTA owns 64 keys, streams query tiles, produces dKf = dS^T Q and dV = P^T dO.
template <int D, int BN, bool KEYS_OUTER>
__global__ void __launch_bounds__(NT) nl_bwd_kernel(int s, int n, int d, const float *__restrict__ q,
                                                    const float *__restrict__ kv, const float *__restrict__ dout,
                                                    const float *__restrict__ lse, const float *__restrict__ delta,
                                                    float scale, float *__restrict__ grad_q,
                                                    float *__restrict__ grad_kv) {
    constexpr int CPT = D / 16, PP = BN + 4, RN = BN / 16;
    extern __shared__ __align__(16) float smem[];
    float *Xat = smem, *Xbt = Xat + D * BM, *Yat = Xbt + D * BM, *Ybt = Yat + D * BN, *Yar = Ybt + D * BN;
    float *dSs = Yar + BN * D;
    float *Ybr = dSs + BM * PP;            // KEYS_OUTER only
    float *Ps = Ybr + BN * D;              // KEYS_OUTER only
    float *lse_c = KEYS_OUTER ? Ps + BM * PP : dSs + BM * PP;  // per-column lse / delta (KEYS_OUTER only)
    float *del_c = lse_c + BN;

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int bi = blockIdx.y, row0 = blockIdx.x * BM;
    const int xcount = KEYS_OUTER ? n : s, ycount = KEYS_OUTER ? s : n;
    const int rows_valid = min(BM, xcount - row0);
    const float *qb = q + (size_t)bi * s * d, *dob = dout + (size_t)bi * s * d, *kvb = kv + (size_t)bi * n * 2 * d;
    const float scale_log2 = scale * kLog2e;

    if constexpr (KEYS_OUTER) {
        load_transposed<D, BM>(Xat, kvb + (size_t)row0 * 2 * d, rows_valid, d, 2 * d);      // Kf
        load_transposed<D, BM>(Xbt, kvb + (size_t)row0 * 2 * d + d, rows_valid, d, 2 * d);  // V
    } else {
        load_transposed<D, BM>(Xat, qb + (size_t)row0 * d, rows_valid, d, d);   // Q
        load_transposed<D, BM>(Xbt, dob + (size_t)row0 * d, rows_valid, d, d);  // dO
    }
    float lse_r[4], del_r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        lse_r[i] = 0.f;
        del_r[i] = 0.f;
        if constexpr (!KEYS_OUTER) {
            const int r = row0 + ty * 4 + i;
            if (r < s) {
                lse_r[i] = __ldg(lse + (size_t)bi * s + r) * kLog2e;
                del_r[i] = __ldg(delta + (size_t)bi * s + r);
            }
        }
    }
    float acc1[4][CPT], acc2[4][CPT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc1[i][c] = acc2[i][c] = 0.f;

    for (int col0 = 0; col0 < ycount; col0 += BN) {
        const int cv = min(BN, ycount - col0);
        __syncthreads();
        if constexpr (KEYS_OUTER) {
            load_transposed<D, BN>(Yat, qb + (size_t)col0 * d, cv, d, d);
            load_transposed<D, BN>(Ybt, dob + (size_t)col0 * d, cv, d, d);
            load_rowmajor<D, BN>(Yar, qb + (size_t)col0 * d, cv, d, d);
            load_rowmajor<D, BN>(Ybr, dob + (size_t)col0 * d, cv, d, d);
            if (threadIdx.x < BN) {
                const bool ok = (int)threadIdx.x < cv;
                lse_c[threadIdx.x] = ok ? __ldg(lse + (size_t)bi * s + col0 + threadIdx.x) * kLog2e : 0.f;
                del_c[threadIdx.x] = ok ? __ldg(delta + (size_t)bi * s + col0 + threadIdx.x) : 0.f;
            }
        } else {
            const float *kvt = kvb + (size_t)col0 * 2 * d;
            load_transposed<D, BN>(Yat, kvt, cv, d, 2 * d);
            load_transposed<D, BN>(Ybt, kvt + d, cv, d, 2 * d);
            load_rowmajor<D, BN>(Yar, kvt, cv, d, 2 * d);
        }
        __syncthreads();
        float sc[4][RN], dp[4][RN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) sc[i][j] = dp[i][j] = 0.f;
        tile_dot<D, BN, RN>(sc, Xat, Yat, ty, tx);
        tile_dot<D, BN, RN>(dp, Xbt, Ybt, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float pr[RN], ds[RN];
#pragma unroll
            for (int j = 0; j < RN; ++j) {
                const int cj = tx * RN + j;
                const bool ok = (ty * 4 + i < rows_valid) && (cj < cv);
                const float l2 = KEYS_OUTER ? lse_c[cj] : lse_r[i];
                const float dl = KEYS_OUTER ? del_c[cj] : del_r[i];
                const float p = ok ? exp2f(fmaf(sc[i][j], scale_log2, -l2)) : 0.f;
                pr[j] = p;
                ds[j] = p * (dp[i][j] - dl) * scale;
            }
            if constexpr (RN == 4) {
                *reinterpret_cast<float4 *>(dSs + (ty * 4 + i) * PP + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
                if constexpr (KEYS_OUTER)
                    *reinterpret_cast<float4 *>(Ps + (ty * 4 + i) * PP + tx * 4) =
                        make_float4(pr[0], pr[1], pr[2], pr[3]);
            } else {
                *reinterpret_cast<float2 *>(dSs + (ty * 4 + i) * PP + tx * 2) = make_float2(ds[0], ds[1]);
                if constexpr (KEYS_OUTER)
                    *reinterpret_cast<float2 *>(Ps + (ty * 4 + i) * PP + tx * 2) = make_float2(pr[0], pr[1]);
            }
        }
        __syncthreads();
        tile_pv<D, BN>(acc1, dSs, Yar, ty, tx);
        if constexpr (KEYS_OUTER) tile_pv<D, BN>(acc2, Ps, Ybr, ty, tx);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + ty * 4 + i;
        if (r < xcount) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = out_col<D>(tx, c);
                if (col < d) {
                    if constexpr (KEYS_OUTER) {
                        float *g = grad_kv + ((size_t)bi * n + r) * 2 * d;
                        g[col] = acc1[i][c];
                        g[d + col] = acc2[i][c];
                    } else {
                        grad_q[((size_t)bi * s + r) * d + col] = acc1[i][c];
                    }
                }
            }
        }
    }
}

template <int D>
size_t fwd_smem() { return sizeof(float) * (size_t)(D * BM + D * 64 + 64 * D + BM * 68); }
template <int D, int BN, bool KO>
size_t bwd_smem() {
    size_t f = (size_t)2 * D * BM + 2 * D * BN + BN * D + BM * (BN + 4);
    if (KO) f += (size_t)BN * D + BM * (BN + 4) + 2 * BN;
    return sizeof(float) * f;
}

template <int D>
int launch_fwd(int b, int s, int n, int d, const float *q, const float *kv, float scale, float *out, float *lse,
               cudaStream_t st) {
    auto kern = nl_fwd_kernel<D>;
    const size_t smem = fwd_smem<D>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div_i(s, BM), (unsigned)b);
    kern<<<grid, NT, smem, st>>>(s, n, d, q, kv, scale * kLog2e, out, lse);
    return pasnl_launch_status();
}

template <int D, int BN>
int launch_bwd(int b, int s, int n, int d, const float *q, const float *kv, const float *out, const float *lse,
               float scale, const float *dout, float *gq, float *gkv, float *delta, cudaStream_t st) {
    const long rows = (long)b * s;
    nl_delta_kernel<<<ceil_div_i(rows, 256), 256, 0, st>>>(rows, d, dout, out, delta);
    cudaError_t e;
    {
        auto kern = nl_bwd_kernel<D, BN, false>;
        const size_t smem = bwd_smem<D, BN, false>();
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)ceil_div_i(s, BM), (unsigned)b);
        kern<<<grid, NT, smem, st>>>(s, n, d, q, kv, dout, lse, delta, scale, gq, gkv);
    }
    {
        auto kern = nl_bwd_kernel<D, BN, true>;
        const size_t smem = bwd_smem<D, BN, true>();
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)ceil_div_i(n, BM), (unsigned)b);
        kern<<<grid, NT, smem, st>>>(s, n, d, q, kv, dout, lse, delta, scale, gq, gkv);
    }
    return pasnl_launch_status();
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace

// tcgen05 / TMEM implementation (nonlocal_tc.cu)
int nl_core_fwd_tcgen05(int b, int s, int n, int d, const float *q, const float *kv, float scale, float *out,
                         float *lse, cudaStream_t st);

// tcgen05 backward (nonlocal_bwd_tc.cu): dQ / (dKf, dV) kernels, after the delta pass
int nl_core_bwd_tcgen05(int b, int s, int n, int d, const float *q, const float *kv, const float *lse, const float *delta,
                         float scale, const float *dout, float *gq, float *gkv, cudaStream_t st);

static bool use_tensor_core_bwd(int s, int n, int d) {
    // PASNL_NL_BWD=simt|tc overrides (profiling / A-B measurements)
    static const char *force = getenv("PASNL_NL_BWD");
    if (force && force[0] == 's') return false;
    if (d != 32) return false;
    if (force && force[0] == 't') return true;
    return s >= 64 && n >= 128;
}

static bool use_tensor_core_path(int s, int n, int d) {
    // A 128-query x 64-key x d tile must be worth filling: tiny problems (deep seg layers: s=40, n=80) and widths the
    // tcgen05 kernel is not instantiated for stay on the CUDA-core kernel.  PASNL_NL_FWD=simt|tc overrides (profiling).
    static const char *force = getenv("PASNL_NL_FWD");
    if (force && force[0] == 's') return false;
    const bool shape_ok = (d == 32 || d == 64);
    if (force && force[0] == 't') return shape_ok;
    return shape_ok && s >= 64 && n >= 128;
}

extern "C" int pasnl_nl_core_fwd(int b, int s, int n, int d, const float *q, const float *kv, float scale, float *out,
                                 float *lse, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && s >= 0 && n >= 1 && d >= 4 && d <= 128 && d % 4 == 0 && b <= 65535);
    if ((long)b * s == 0) return PASNL_OK;
    PASNL_CHECK_ARG(q && kv && out && lse && aligned16(q) && aligned16(kv));
    cudaStream_t st = (cudaStream_t)stream;
    if (use_tensor_core_path(s, n, d) && aligned16(out)) return nl_core_fwd_tcgen05(b, s, n, d, q, kv, scale, out, lse, st);
    if (d <= 32) return launch_fwd<32>(b, s, n, d, q, kv, scale, out, lse, st);
    if (d <= 64) return launch_fwd<64>(b, s, n, d, q, kv, scale, out, lse, st);
    return launch_fwd<128>(b, s, n, d, q, kv, scale, out, lse, st);
}

extern "C" size_t pasnl_nl_core_bwd_workspace_bytes(int b, int s, int n, int d) {
    (void)n;
    (void)d;
    if (b <= 0 || s <= 0) return 0;
    return sizeof(float) * (size_t)b * (size_t)s;  // delta
}

extern "C" int pasnl_nl_core_bwd(int b, int s, int n, int d, const float *q, const float *kv, const float *out,
                                 const float *lse, float scale, const float *grad_out, float *grad_q, float *grad_kv,
                                 void *workspace, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && s >= 0 && n >= 1 && d >= 4 && d <= 128 && d % 4 == 0 && b <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    if ((long)b * s == 0) {
        if (b > 0 && grad_kv) {
            cudaError_t e = cudaMemsetAsync(grad_kv, 0, sizeof(float) * (size_t)b * n * 2 * d, st);
            if (e != cudaSuccess) return (int)e;
        }
        return PASNL_OK;
    }
    PASNL_CHECK_ARG(q && kv && out && lse && grad_out && grad_q && grad_kv && workspace);
    PASNL_CHECK_ARG(aligned16(q) && aligned16(kv) && aligned16(out) && aligned16(grad_out));
    float *delta = (float *)workspace;
    if (use_tensor_core_bwd(s, n, d) && aligned16(grad_q) && aligned16(grad_kv)) {
        const long rows = (long)b * s;
        nl_delta_kernel<<<ceil_div_i(rows, 256), 256, 0, st>>>(rows, d, grad_out, out, delta);
        return nl_core_bwd_tcgen05(b, s, n, d, q, kv, lse, delta, scale, grad_out, grad_q, grad_kv, st);
    }
    if (d <= 32) return launch_bwd<32, 64>(b, s, n, d, q, kv, out, lse, scale, grad_out, grad_q, grad_kv, delta, st);
    if (d <= 64) return launch_bwd<64, 64>(b, s, n, d, q, kv, out, lse, scale, grad_out, grad_q, grad_kv, delta, st);
    return launch_bwd<128, 32>(b, s, n, d, q, kv, out, lse, scale, grad_out, grad_q, grad_kv, delta, st);
}
