// gemm_tma.cu -- TMA-fed rows GEMM for the MID-SIZE layers (rows 4096..65536, k or n beyond one TMEM-resident weight
// tile), sm_100a:   C[rows, n] = A[rows, k] . Wt[n, k]^T,   3xTF32.
//
// Same orientation as gemm_ws.cu (C^T = Wt . A^T: TMEM lane = output channel, so the epilogue stores coalesced rows
// straight from registers and the batch-norm column statistics are thread-local), but the weights do not fit TMEM
// here, so BOTH operands arrive by TMA tensor copies as 128 x 32 SWIZZLE_128B K-major tiles: per stage the TMA warp
// loads a Wt tile (UMMA A operand, SS form) and an activation tile (UMMA B operand); the maker warps derive the lo
// parts of both in shared memory; the MMA warp issues 4 k-steps x (Whi.Xhi, Wlo.Xhi, Whi.Xlo).  Persistent CTAs walk
// (row tile, channel tile) pairs with the channel tile fastest, so neighbouring CTAs share their activation rows in L2.
// Contractions longer than 1024 are accumulated in <= 1024-wide slabs (the TMEM adder truncates, ~3e-8*k relative):
// the slabs of a tile alternate between the two accumulators and the epilogue adds them in fp32 (read-modify-write of
// its own rows of C: no atomics, no memset), so the column statistics stay fused for long k too.
// This replaces gemm_tc.cu's one-tile kernel (register-staged loads, 2 x 8 KB stages, one __syncthreads per 16-wide
// k chunk: 63..83 TFLOP/s on these layers) wherever the operands are TMA-legal (16-byte aligned, k % 4 == 0).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

using namespace pasnl;

namespace gtma {

constexpr int kRows = 128;                       // rows per tile = UMMA N
constexpr int kKB = 32;                          // k per stage: 32 floats = the 128-byte swizzle span
constexpr int kTileBytes = kRows * kKB * 4;      // 16 KB (hi) ; a stage is hi | lo
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kStages = 3;                       // x (W hi | W lo | X hi | X lo) = 64 KB
constexpr int kFullStage = 2 * kStageBytes;
constexpr int kThreads = 14 * 32;
constexpr int kTmaWarp = 4, kMmaWarp = 5, kMakerWarp0 = 6, kMakerWarps = 8;
constexpr int kAccCols = 128;                    // TMEM columns per accumulator (one per tile row)
constexpr int kSlabChunks = 32;                  // 32 x 32 = 1024 contraction columns per accumulation (TMEM adder truncates)

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

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_rows_tma_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap, long rows, int k,
                     int n, int n_tiles, long ntiles, float *__restrict__ C, long ldc, double *__restrict__ stats) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t loaded[kStages], ready[kStages], empty[kStages], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t *stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    if (t == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&loaded[s], 1); mbar_init(&ready[s], kMakerWarps); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == kTmaWarp && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int nchunks = (k + kKB - 1) / kKB;
    const int nslabs = (nchunks + kSlabChunks - 1) / kSlabChunks;
    const long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int s = 0;
            uint32_t par = 1;
            for (long j = 0; j < my_tiles; ++j) {
                const long tile = (long)blockIdx.x + j * gridDim.x;
                const int row0 = (int)(tile / n_tiles) * kRows, n0 = (int)(tile % n_tiles) * 128;
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(&empty[s], par);
                    mbar_expect_tx(&loaded[s], 2 * kTileBytes);
                    uint8_t *st = stages + (size_t)s * kFullStage;
                    tma_load_2d(st, &wmap, c * kKB, n0, &loaded[s]);
                    tma_load_2d(st + kStageBytes, &amap, c * kKB, row0, &loaded[s]);
                    if (++s == kStages) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        constexpr uint32_t idesc = make_idesc(128, kRows);
        int s = 0;
        uint32_t par = 0;
        long item = 0;                                            // (tile, slab) counter: accumulator = item & 1
        for (long j = 0; j < my_tiles; ++j) {
            for (int sl = 0; sl < nslabs; ++sl, ++item) {
                const int b = (int)(item & 1);
                mbar_wait(&tempty[b], (uint32_t)(((item >> 1) & 1) ^ 1));
                tc_fence_after();
                const int c0 = sl * kSlabChunks, c1 = min(nchunks, c0 + kSlabChunks);
                for (int c = c0; c < c1; ++c) {
                    mbar_wait(&ready[s], par);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w_hi = smem_u32(stages + (size_t)s * kFullStage), w_lo = w_hi + kTileBytes;
                        const uint32_t x_hi = w_hi + kStageBytes, x_lo = x_hi + kTileBytes;
                        const uint32_t d = tmem + b * kAccCols;
#pragma unroll
                        for (int ks = 0; ks < kKB / 8; ++ks) {
                            const uint64_t wh = make_desc_sw128(w_hi + ks * 32), wl = make_desc_sw128(w_lo + ks * 32);
                            const uint64_t xh = make_desc_sw128(x_hi + ks * 32), xl = make_desc_sw128(x_lo + ks * 32);
                            mma_tf32_ss(d, wh, xh, idesc, (c > c0 || ks > 0) ? 1u : 0u);
                            mma_tf32_ss(d, wl, xh, idesc, 1);
                            mma_tf32_ss(d, wh, xl, idesc, 1);
                        }
                        mma_commit(&empty[s]);
                        if (c == c1 - 1) mma_commit(&tfull[b]);
                    }
                    __syncwarp();
                    if (++s == kStages) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp >= kMakerWarp0) {
        // makers: tf32(x) in place, x - tf32(x) beside it, for the weight tile and the activation tile of the stage
        const int mt = t - kMakerWarp0 * 32;
        int s = 0;
        uint32_t par = 0;
        for (long it = 0; it < my_tiles * nchunks; ++it) {
            mbar_wait(&loaded[s], par);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float4 *hi = reinterpret_cast<float4 *>(stages + (size_t)s * kFullStage + half * kStageBytes) + mt;
                float4 *lo = hi + kTileBytes / 16;
                float4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = hi[256 * i];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = x[i];
                    float4 h, l;
                    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    hi[256 * i] = h;
                    lo[256 * i] = l;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
            if (++s == kStages) { s = 0; par ^= 1u; }
        }
    } else {
        // epilogue: lane = output channel of this tile's 128-channel window
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        long item = 0;
        for (long j = 0; j < my_tiles; ++j) {
            const long tile = (long)blockIdx.x + j * gridDim.x;
            const long row0 = (tile / n_tiles) * kRows;
            const int ch = (int)(tile % n_tiles) * 128 + warp * 32 + lane;
            const bool chv = ch < n;
            const int nv = (int)min((long)kRows, rows - row0);
            for (int sl = 0; sl < nslabs; ++sl, ++item) {
                const int b = (int)(item & 1);
                const bool last = sl == nslabs - 1;
                mbar_wait(&tfull[b], (uint32_t)((item >> 1) & 1));
                tc_fence_after();
                float p1 = 0.f, p2 = 0.f, shiftv = 0.f;
#pragma unroll 1
                for (int q = 0; q < kRows / 32; ++q) {
                    float v[32];
                    tmem_ld32(tmem + lane_sel + b * kAccCols + q * 32, v);
                    if (q == kRows / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[b]);
                    }
                    float *dst = C + (row0 + q * 32) * ldc + ch;
                    if (chv) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (q * 32 + i < nv) {
                                float x = v[i];
                                if (sl > 0) x += dst[(long)i * ldc];      // earlier slabs of this tile, written by this thread
                                dst[(long)i * ldc] = x;
                                if (last) {
                                    if (q == 0 && i == 0) shiftv = x;
                                    const float dlt = x - shiftv;
                                    p1 += dlt;
                                    p2 = fmaf(dlt, dlt, p2);
                                }
                            }
                        }
                    }
                }
                if (last && stats != nullptr && chv && nv > 0) {
                    const double kd = (double)shiftv, d1 = (double)p1, c = (double)nv;
                    atomicAdd(stats + ch, d1 + c * kd);
                    atomicAdd(stats + n + ch, (double)p2 + 2.0 * kd * d1 + c * kd * kd);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
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


}  // namespace gtma

// Whether pasnl_gemm_rows_tf32x3 may route a problem here.  PASNL_GEMM_TMA=0 disables the kernel (A/B measurements).
extern "C" int pasnl_gemm_rows_tma_eligible(long rows, int k, int n, const void *A, long lda, const void *Wt, long ldw) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_GEMM_TMA");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || gtma::encode_fn() == nullptr) return 0;
    if ((lda % 4) != 0 || (ldw % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15u) != 0 ||
        (reinterpret_cast<uintptr_t>(Wt) & 15u) != 0)
        return 0;
    if (n < 64 || rows < 2048 || rows >= 2147483647L - 128) return 0;
    // measured on B200 (tools/sweep_small_gemm.py, profiles/r02_sweep_mid_gemm.json): the TMA pipeline pays off for long
    // contractions with enough (row tile, channel tile) pairs to cover the SMs -- 16384x1024x128: 45 vs 66 us,
    // 4096x1024x512: 45 vs 64 us; short k or few tiles stay on the one-tile kernel (its split-k fills the GPU better:
    // 4096x4096x256: 99 vs 170 us; 16384x128x2048: 113 vs 141 us)
    const long tiles = ((rows + 127) / 128) * ((n + 127) / 128);
    if (k < 512 || tiles < (long)gtma::sm_count() * 2 / 3) return 0;
    return 1;
}

// C[rows, n] = A[rows, k] . Wt[n, k]^T; col_stats (2n doubles, zero-initialised) or NULL: += column sum / sum of squares.
extern "C" int pasnl_gemm_rows_tma(long rows, int k, int n, const float *A, long lda, const float *Wt, long ldw, float *C,
                                   long ldc, double *col_stats, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && k >= 1 && n >= 1 && lda >= k && ldw >= k && ldc >= n);
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(A && Wt && C && pasnl_gemm_rows_tma_eligible(rows, k, n, A, lda, Wt, ldw));
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap amap, wmap;
    const cuuint32_t box[2] = {(cuuint32_t)gtma::kKB, (cuuint32_t)gtma::kRows};
    const cuuint32_t estr[2] = {1, 1};
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)lda * 4};
        if (gtma::encode_fn()(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)A, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return PASNL_EINVAL;
    }
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)n};
        const cuuint64_t gstride[1] = {(cuuint64_t)ldw * 4};
        if (gtma::encode_fn()(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)Wt, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return PASNL_EINVAL;
    }
    const int n_tiles = (n + 127) / 128;
    const long ntiles = ((rows + gtma::kRows - 1) / gtma::kRows) * n_tiles;
    const int grid = (int)(ntiles < gtma::sm_count() ? ntiles : gtma::sm_count());
    const size_t smem = (size_t)gtma::kStages * gtma::kFullStage + 1024;
    cudaError_t e = cudaFuncSetAttribute(gtma::gemm_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gtma::gemm_rows_tma_kernel<<<grid, gtma::kThreads, smem, st>>>(amap, wmap, rows, k, n, n_tiles, ntiles, C, ldc, col_stats);
    return pasnl_launch_status();
}
