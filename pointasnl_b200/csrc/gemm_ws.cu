// gemm_ws.cu -- rows GEMM with the WEIGHTS STATIONARY IN TMEM and the activations streamed by TMA, sm_100a.
//
//     C[rows, n] = f(A)[rows, k] . W          rows ~ 10^5..10^6,  k <= 128,  n <= 128,  3xTF32 (fp32-accurate)
//     f = identity, or a per-input-column affine + activation (the batch-norm normalise + ReLU of the layer that
//     produced A, folded into this kernel's operand path so that the normalised tensor never exists in HBM)
//
// Why a second rows kernel.  ncu on gemm_tc.cu's persistent kernel (profiles/r01b_*): the L1 / shared-memory data pipe
// is 80 % busy -- 6.3 M wavefronts of tensor-core operand reads (3 MMAs per k-step each re-read A and B from shared
// memory), 6.2 M of operand STS + epilogue panel traffic -- and its producers stall on global loads they issue
// themselves (48 KB in flight per SM).  This kernel turns the product around, C^T = W^T . f(A)^T:
//   * UMMA "A" operand = W^T (M = n output channels x K): split once into hi/lo TF32 parts and kept IN TMEM for the
//     whole launch (tcgen05.st, TS-form tcgen05.mma) -- the weights cost no shared memory and no operand reads;
//   * UMMA "B" operand = a 128-row x 32-column activation tile, K-major, SWIZZLE_128B, written into a 6-stage ring by
//     TMA tensor copies (cp.async.bulk.tensor.2d -> SASS UTMALDG): 96 KB in flight per SM, no address arithmetic,
//     no register staging; 8 "maker" warps only derive the lo part (x - tf32(x)) tile-locally, smem -> smem;
//   * the accumulator is C^T (TMEM lane = output channel, column = row): an epilogue thread owns ONE channel, so the
//     batch-norm column statistics are thread-local sums, and 32 lanes = 32 consecutive channels of one row = one
//     coalesced 128-byte store straight from registers -- no shared-memory panel.
// Shared-memory traffic per 128x128x128 tile: 64 KB TMA write + 192 KB maker + 192 KB operand reads (was 640 KB).
//
// Warp roles (14 warps, one CTA per SM, persistent over row tiles):
//   0..3   epilogue (TMEM quadrant w): load W into TMEM at start; then drain accumulators, statistics, global stores
//   4      TMA producer (one elected lane)
//   5      TMEM allocation + MMA issue (one elected lane): per stage 4 k-steps x (Whi.Xhi, Wlo.Xhi, Whi.Xlo)
//   6..13  makers: wait for the TMA bytes, apply f, write tf32(x) in place and the remainder to the lo tile
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

using namespace pasnl;

namespace gws {

constexpr int kRows = 128;                       // rows per tile = UMMA N
constexpr int kKB = 32;                          // k per stage: 32 floats = the 128-byte swizzle span
constexpr int kTileBytes = kRows * kKB * 4;      // 16 KB (hi) ; a stage is hi | lo
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kStages = 6;
constexpr int kThreads = 14 * 32;
constexpr int kTmaWarp = 4, kMmaWarp = 5, kMakerWarp0 = 6, kMakerWarps = 8;
constexpr int kAccCols = 128;                    // TMEM columns per accumulator (one per tile row)
constexpr int kWCol = 2 * kAccCols;              // W hi at [256, 256+kpad), W lo at [256+kpad, 256+2*kpad)

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {   // kind::tf32, fp32 accumulate, K-major A and B
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// K-major SWIZZLE_128B operand: 8-row groups 1024 bytes apart (SBO), LBO unused, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7]))
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// w_kn != 0: W is stored [k][n] (row pitch ldw), the layer's own weight layout; else [n][k] (the transposed view the
// input-gradient product uses: dx = dy . W^T reads W [cin][cout] as "n = cin rows of k = cout").
__global__ void __launch_bounds__(kThreads, 1)
gemm_rows_ws_kernel(const __grid_constant__ CUtensorMap amap, long rows, int k, int kpad, int n, long ntiles,
                    const float *__restrict__ W, long ldw, int w_kn, float *__restrict__ C, long ldc,
                    double *__restrict__ stats, const float *__restrict__ pro_scale, const float *__restrict__ pro_shift,
                    int pro_act) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t loaded[kStages], ready[kStages], empty[kStages], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_scale[128], s_shift[128];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t *stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned

    if (t == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&loaded[s], 1); mbar_init(&ready[s], kMakerWarps); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == kTmaWarp && lane == 0)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
    if (t < 128) {
        const bool has = pro_scale != nullptr && t < k;
        s_scale[t] = has ? __ldg(pro_scale + t) : (t < k ? 1.f : 0.f);
        s_shift[t] = has ? __ldg(pro_shift + t) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // blockIdx.y = 128-column block of the output (n > 128: the "wide" launch, one weight block per CTA)
    const int n0 = blockIdx.y * 128, n_here = min(128, n - n0);
    if (warp < 4) {   // W^T -> TMEM, split once: lane = output channel, column = k
        const int ch = warp * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        for (int k0 = 0; k0 < kpad; k0 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = k0 + i;
                float w = 0.f;
                if (ch < n_here && kk < k) w = w_kn ? __ldg(W + (long)kk * ldw + n0 + ch) : __ldg(W + (long)(n0 + ch) * ldw + kk);
                hi[i] = tf32_hi(w);
                lo[i] = w - hi[i];
            }
            tmem_st8(tmem + lane_sel + kWCol + k0, hi);
            tmem_st8(tmem + lane_sel + kWCol + kpad + k0, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int nchunks = kpad / kKB;
    const long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

    if (warp == kTmaWarp) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t par = 1;                                     // empty[] starts "already released"
            for (long j = 0; j < my_tiles; ++j) {
                const long tile = (long)blockIdx.x + j * gridDim.x;
                const int row0 = (int)(tile * kRows);
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(&empty[s], par);
                    mbar_expect_tx(&loaded[s], kTileBytes);
                    tma_load_2d(stages + (size_t)s * kStageBytes, &amap, c * kKB, row0, &loaded[s]);
                    if (++s == kStages) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = make_idesc(128, kRows);
        int s = 0;
        uint32_t par = 0;
        for (long j = 0; j < my_tiles; ++j) {
            const int b = (int)(j & 1);
            mbar_wait(&tempty[b], (uint32_t)(((j >> 1) & 1) ^ 1));
            tc_fence_after();
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&ready[s], par);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t x_hi = smem_u32(stages + (size_t)s * kStageBytes), x_lo = x_hi + kTileBytes;
                    const uint32_t w_hi = tmem + kWCol + c * kKB, w_lo = w_hi + kpad;
                    const uint32_t d = tmem + b * kAccCols;
#pragma unroll
                    for (int ks = 0; ks < kKB / 8; ++ks) {
                        const uint64_t xh = make_desc_sw128(x_hi + ks * 32);
                        const uint64_t xl = make_desc_sw128(x_lo + ks * 32);
                        mma_tf32_ts(d, w_hi + ks * 8, xh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32_ts(d, w_lo + ks * 8, xh, idesc, 1);
                        mma_tf32_ts(d, w_hi + ks * 8, xl, idesc, 1);
                    }
                    mma_commit(&empty[s]);
                    if (c == nchunks - 1) mma_commit(&tfull[b]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; par ^= 1u; }
            }
        }
    } else if (warp >= kMakerWarp0) {
        // ------------------------------------------------------------------ makers: hi in place, lo beside it
        // thread mt owns the 16-byte pieces mt + 256 i (i = 0..3) of every tile: physical chunk mt % 8 of rows
        // mt / 8 + 32 i.  Under the 128-byte swizzle that is LOGICAL chunk (mt % 8) ^ (row % 8), and row % 8 does not
        // depend on i -> one logical k offset per thread: its scale / shift are the same 4 columns for all 4 pieces.
        const int mt = t - kMakerWarp0 * 32;
        const int jl = (mt & 7) ^ ((mt >> 3) & 7);
        const bool pro = pro_scale != nullptr;
        int s = 0;
        uint32_t par = 0;
        for (long j = 0; j < my_tiles; ++j) {
            for (int c = 0; c < nchunks; ++c) {
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pro) {
                    sc = *reinterpret_cast<const float4 *>(s_scale + c * kKB + jl * 4);
                    sh = *reinterpret_cast<const float4 *>(s_shift + c * kKB + jl * 4);
                }
                mbar_wait(&loaded[s], par);
                float4 *hi = reinterpret_cast<float4 *>(stages + (size_t)s * kStageBytes) + mt;
                float4 *lo = hi + kTileBytes / 16;
                float4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = hi[256 * i];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 v = x[i];
                    if (pro) {
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                        if (pro_act == 1) {
                            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                        } else if (pro_act == 2) {
                            v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y;
                            v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w;
                        }
                    }
                    float4 h, l;
                    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    hi[256 * i] = h;
                    lo[256 * i] = l;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[s]);
                if (++s == kStages) { s = 0; par ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: C^T rows = this thread's channel
        const int ch = warp * 32 + lane;
        const bool chv = ch < n_here;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        double S1 = 0.0, S2 = 0.0;
        long cnt = 0;
        float shiftv = 0.f;
        bool have_shift = false;
        for (long j = 0; j < my_tiles; ++j) {
            const int b = (int)(j & 1);
            const long row0 = ((long)blockIdx.x + j * gridDim.x) * kRows;
            const int nv = (int)min((long)kRows, rows - row0);
            mbar_wait(&tfull[b], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            float p1 = 0.f, p2 = 0.f;
#pragma unroll
            for (int q = 0; q < kRows / 32; ++q) {
                float v[32];
                tmem_ld32(tmem + lane_sel + b * kAccCols + q * 32, v);
                if (q == kRows / 32 - 1) {                        // accumulator fully read: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[b]);
                }
                if (!have_shift) { shiftv = v[0]; have_shift = true; }
                float *dst = C + (row0 + q * 32) * ldc + n0 + ch;
                if (chv) {
                    if (q * 32 + 32 <= nv) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            dst[(long)i * ldc] = v[i];
                            const float dlt = v[i] - shiftv;
                            p1 += dlt;
                            p2 = fmaf(dlt, dlt, p2);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (q * 32 + i < nv) {
                                dst[(long)i * ldc] = v[i];
                                const float dlt = v[i] - shiftv;
                                p1 += dlt;
                                p2 = fmaf(dlt, dlt, p2);
                            }
                        }
                    }
                }
            }
            S1 += (double)p1;
            S2 += (double)p2;
            cnt += nv;
        }
        if (stats != nullptr && chv && cnt > 0) {
            const double kd = (double)shiftv, c = (double)cnt;
            atomicAdd(stats + n0 + ch, S1 + c * kd);
            atomicAdd(stats + n + n0 + ch, S2 + 2.0 * kd * S1 + c * kd * kd);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

inline int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace gws

// Whether pasnl_gemm_rows_tf32x3 may route a problem here: persistent regime, TMA-legal A (16-byte aligned base and row
// pitch), k and n within one TMEM-resident weight tile.  PASNL_GEMM_WS=0 disables the kernel (A/B measurements).
extern "C" int pasnl_gemm_rows_ws_eligible(long rows, int k, int n, const void *A, long lda) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_GEMM_WS");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || gws::encode_fn() == nullptr) return 0;
    if (k < 1 || k > 128 || n < 1 || n > 128) return 0;
    // measured on B200 (tools/bench_gemm.py, profiles/r02_gemm_ws_vs_persist.json): the 32-wide TMA box and the M = 128
    // weight tile are mostly padding for narrow layers (4194304x16x16: 328 vs 267 us, 1310720x32x32: 109 vs 95 us)
    if (!((k > 16 && n > 32) || (k > 32 && n > 16))) return 0;
    if ((lda % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15u) != 0) return 0;
    if (rows < 2L * gws::kRows * gws::sm_count() || rows / gws::kRows >= 2147483647L / gws::kRows) return 0;
    return 1;
}

// n > 128 ("wide"): the same kernel with one 128-column weight block per CTA (blockIdx.y) and the row tiles of a block shared
// by sm_count / blocks persistent CTAs; A (k <= 128 columns) is re-read per block out of L2.  For products that are bound
// by WRITING C (the input gradient of the [1,K] after_conv layer: 16384 x 128 -> 2048, 134 MB out) the coalesced
// C^T-epilogue stores are what matters.  Eligible when there are enough (row tile, block) pairs to amortise the per-CTA
// weight split: >= 4 per SM.
extern "C" int pasnl_gemm_rows_ws_wide_eligible(long rows, int k, int n, const void *A, long lda) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_GEMM_WS_WIDE");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || gws::encode_fn() == nullptr) return 0;
    // (measured on the cls step: the 134-wide input gradient of cls layer 2 as ONE wide launch -- a second column block
    //  of persistent CTAs on 6 useful columns -- beats 128 columns here + the remainder through the one-tile kernel)
    if (k <= 32 || k > 128 || n <= 128 || n > 128 * 64) return 0;
    if ((lda % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15u) != 0) return 0;
    const long ntiles = (rows + gws::kRows - 1) / gws::kRows, nblocks = (n + 127) / 128;
    if (nblocks > gws::sm_count() || ntiles * nblocks < 4L * gws::sm_count() || ntiles >= 2147483647L / gws::kRows) return 0;
    return 1;
}

// C[rows, n] = f(A)[rows, k] . W.  W is [k][n] (w_kn != 0, pitch ldw) or [n][k] (w_kn == 0).  col_stats (2n doubles,
// zero-initialised by the caller) or NULL: += per-column sum and sum of squares of C.  pro_scale / pro_shift (k floats
// each) or NULL: f(a) = act(a * scale[col] + shift[col]), pro_act 0 none / 1 relu / 2 leaky_relu(0.2).
extern "C" int pasnl_gemm_rows_ws(long rows, int k, int n, const float *A, long lda, const float *W, long ldw, int w_kn,
                                  float *C, long ldc, double *col_stats, const float *pro_scale, const float *pro_shift,
                                  int pro_act, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && k >= 1 && n >= 1 && lda >= k && ldc >= n && (w_kn ? ldw >= n : ldw >= k));
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(A && W && C && (pasnl_gemm_rows_ws_eligible(rows, k, n, A, lda) ||
                                    pasnl_gemm_rows_ws_wide_eligible(rows, k, n, A, lda)));
    PASNL_CHECK_ARG((pro_scale == nullptr) == (pro_shift == nullptr) && pro_act >= 0 && pro_act <= 2);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)lda * 4};
    const cuuint32_t box[2] = {(cuuint32_t)gws::kKB, (cuuint32_t)gws::kRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = gws::encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)A, gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PASNL_EINVAL;
    const int kpad = (k + gws::kKB - 1) / gws::kKB * gws::kKB;
    const long ntiles = (rows + gws::kRows - 1) / gws::kRows;
    const int nblocks = (n + 127) / 128;
    long gx = gws::sm_count() / nblocks;                       // persistent CTAs per column block
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    const dim3 grid((unsigned)gx, (unsigned)nblocks);
    const size_t smem = (size_t)gws::kStages * gws::kStageBytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(gws::gemm_rows_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gws::gemm_rows_ws_kernel<<<grid, gws::kThreads, smem, st>>>(map, rows, k, kpad, n, ntiles, W, ldw, w_kn, C, ldc,
                                                                col_stats, pro_scale, pro_shift, pro_act);
    return pasnl_launch_status();
}
