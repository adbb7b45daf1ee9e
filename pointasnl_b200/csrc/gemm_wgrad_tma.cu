// gemm_wgrad_tma.cu -- weight-gradient GEMM with BOTH operands streamed by TMA as MN-major tiles, sm_100a.
//
//     dW[m, n] = sum_r X[r, m] * Y[r, n]        rows ~ 10^4..10^6 (the contraction), m, n = layer widths, 3xTF32
//
// X and Y arrive row-major with the contraction index r as the SLOW one, i.e. both are "MN-major" UMMA operands.  The
// round-1 kernels transposed them on the way into shared memory (shuffle / select chains in gemm_tc.cu's Q producers,
// the hot spot of profiles/r01b_ncu_source_hotspots_gemm_wgrad_tm.txt) or fed one of them through TMEM.  tcgen05 reads
// MN-major tf32 operands directly when the tile uses the 128-byte swizzle with 32-byte atomicity
// (UMMA layout type SWIZZLE_128B_BASE32B), which is exactly what a TMA tensor copy with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes: a box of 32 columns x 32 rows lands as 32 rows of 128 bytes, 4-row atoms
// of 512 bytes.  So here nothing is transposed and no thread touches an address:
//   TMA warp    per stage (32 contraction rows): 4 boxes of X (128 columns), n/32 boxes of Y  -> 16 + 16 KB, into a
//               4-deep RAW ring (three stages of loads in flight: the tall layers are bound by load latency)
//   makers      tf32(x) in place, lo = x - tf32(x) into a 2-deep lo ring, element-wise, layout-agnostic
//   MMA warp    4 k-steps (8 rows) x (Xhi.Yhi, Xlo.Yhi, Xhi.Ylo); descriptors: start + 1024 B per k-step,
//               LBO = 4096 B between 32-column chunks, SBO = 512 B between 4-row atoms, a_major = b_major = MN
//   epilogue    every 1024 contraction rows the accumulator (the TMEM adder truncates) is added into dW with red.add,
//               alternating between two accumulators so that the tensor core never waits for the drain
// One CTA per (row slab, 128-wide m tile, <= 128-wide n tile).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

using namespace pasnl;

namespace gwt {

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


constexpr int kKRows = 32;                       // contraction rows per stage
constexpr int kChunkBytes = kKRows * 128;        // one 32-column x 32-row box: 4 KB
constexpr int kOpBytes = 4 * kChunkBytes;        // 128 columns: 16 KB
constexpr int kRawStage = 2 * kOpBytes;          // X | Y as the TMA writes them (tf32(x) in place after the makers): 32 KB
constexpr int kRaw = 4;                          // raw ring: up to 3 stages of loads in flight while one is consumed
constexpr int kLo = 2;                           // lo ring (X lo | Y lo, 32 KB): only stages in the maker / MMA phase need one
constexpr int kSubStages = 32;                   // 32 x 32 = 1024 rows per accumulation

__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {   // kind::tf32, fp32 accumulate, MN-major A and B
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// MN-major SWIZZLE_128B_BASE32B operand: 32-column chunks LBO bytes apart, 4-row atoms SBO bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((uint32_t)kChunkBytes >> 4) << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_tma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap, long rows, long slab,
                      int m, int n, int nq, float *__restrict__ dW, long ldw, const float *__restrict__ x_scale,
                      const float *__restrict__ x_shift, int x_act) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t loaded[kRaw], empty[kRaw], ready[kLo], lofree[kLo], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_scale[128], s_shift[128];   // operand prologue of this CTA's 128 X columns
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t *stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *lobuf = stages + (size_t)kRaw * kRawStage;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.z * nq;
    const long r_begin = (long)blockIdx.x * slab, r_end = min(rows, r_begin + slab);
    if (r_begin >= r_end) return;
    const int nst = (int)((r_end - r_begin + kKRows - 1) / kKRows);      // stages of 32 rows (TMA zero-fills past `rows`)
    const int nsub = (nst + kSubStages - 1) / kSubStages;
    const int ychunks = nq / 32;
    const bool pro = x_scale != nullptr;
    if (pro && t < 128) {
        const int col = m0 + t;
        s_scale[t] = col < m ? __ldg(x_scale + col) : 0.f;
        s_shift[t] = col < m ? __ldg(x_shift + col) : 0.f;
    }

    if (t == 0) {
        for (int s = 0; s < kRaw; ++s) { mbar_init(&loaded[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < kLo; ++s) { mbar_init(&ready[s], kMakerWarps); mbar_init(&lofree[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int s = 0;
            uint32_t par = 1;
            for (int i = 0; i < nst; ++i) {
                const int r0 = (int)(r_begin + (long)i * kKRows);
                mbar_wait(&empty[s], par);
                mbar_expect_tx(&loaded[s], (uint32_t)(4 + ychunks) * kChunkBytes);
                uint8_t *st = stages + (size_t)s * kRawStage;
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_2d(st + c * kChunkBytes, &xmap, m0 + 32 * c, r0, &loaded[s]);
                for (int c = 0; c < ychunks; ++c) tma_load_2d(st + kOpBytes + c * kChunkBytes, &ymap, n0 + 32 * c, r0, &loaded[s]);
                if (++s == kRaw) { s = 0; par ^= 1u; }
            }
        }
    } else if (warp == kMmaWarp) {
        const uint32_t idesc = make_idesc_mn(128, nq);
        int s = 0;
        for (int i = 0; i < nst; ++i) {
            const int sub = i / kSubStages, ci = i - sub * kSubStages, b = sub & 1, l = i & 1;
            if (ci == 0) {
                mbar_wait(&tempty[b], (uint32_t)(((sub >> 1) & 1) ^ 1));
                tc_fence_after();
            }
            mbar_wait(&ready[l], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            if (elect_one()) {
                const uint32_t x_hi = smem_u32(stages + (size_t)s * kRawStage), y_hi = x_hi + kOpBytes;
                const uint32_t x_lo = smem_u32(lobuf + (size_t)l * kRawStage), y_lo = x_lo + kOpBytes;
                const uint32_t d = tmem + b * 128;
#pragma unroll
                for (int ks = 0; ks < kKRows / 8; ++ks) {
                    const uint64_t xh = make_desc_mn(x_hi + ks * 1024), xl = make_desc_mn(x_lo + ks * 1024);
                    const uint64_t yh = make_desc_mn(y_hi + ks * 1024), yl = make_desc_mn(y_lo + ks * 1024);
                    mma_tf32_ss(d, xh, yh, idesc, (ci > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32_ss(d, xl, yh, idesc, 1);
                    mma_tf32_ss(d, xh, yl, idesc, 1);
                }
                mma_commit(&empty[s]);
                mma_commit(&lofree[l]);
                if (ci == kSubStages - 1 || i == nst - 1) mma_commit(&tfull[b]);
            }
            __syncwarp();
            if (++s == kRaw) s = 0;
        }
    } else if (warp >= kMakerWarp0) {
        const int mt = t - kMakerWarp0 * 32;
        int s = 0;
        uint32_t par = 0;
        for (int i = 0; i < nst; ++i) {
            const int l = i & 1;
            mbar_wait(&loaded[s], par);
            mbar_wait(&lofree[l], (uint32_t)(((i >> 1) & 1) ^ 1));         // lo slot l was read by the MMAs of stage i - 2
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float4 *hi = reinterpret_cast<float4 *>(stages + (size_t)s * kRawStage + half * kOpBytes) + mt;
                float4 *lo = reinterpret_cast<float4 *>(lobuf + (size_t)l * kRawStage + half * kOpBytes) + mt;
                const int pieces = half == 0 ? 4 : ychunks;            // 256 threads x 16 B = one 4 KB chunk per pass
                for (int p = 0; p < pieces; ++p) {
                    float4 v = hi[256 * p];
                    if (pro && half == 0) {
                        // X <- act(X * scale[col] + shift[col]) (the producing layer's batch norm + activation).  Thread mt
                        // holds row mt/8 of the chunk and the 16-byte piece mt%8 of its 128-byte line; the 32-byte atoms
                        // of a line are XOR-swizzled with (row & 3): logical column = ((piece>>1) ^ (row&3))*8 + (piece&1)*4
                        const int row = mt >> 3, piece = mt & 7;
                        const int col = p * 32 + (((piece >> 1) ^ (row & 3)) << 3) + ((piece & 1) << 2);
                        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + col);
                        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + col);
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                        if (x_act == 1) {
                            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                        } else if (x_act == 2) {
                            v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y;
                            v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w;
                        }
                        if (r_begin + (long)i * kKRows + row >= r_end) v = make_float4(0.f, 0.f, 0.f, 0.f);   // rows TMA zero-filled
                    }
                    float4 h, l;
                    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    hi[256 * p] = h;
                    lo[256 * p] = l;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[l]);
            if (++s == kRaw) { s = 0; par ^= 1u; }
        }
    } else {
        // epilogue: lane = row of dW (m index): add finished 1024-row partial sums into dW
        const int mrow = m0 + warp * 32 + lane;
        const bool valid = mrow < m;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        for (int sub = 0; sub < nsub; ++sub) {
            const int b = sub & 1;
            mbar_wait(&tfull[b], (uint32_t)((sub >> 1) & 1));
            tc_fence_after();
            for (int c0 = 0; c0 < nq; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + lane_sel + b * 128 + c0, v);
                if (valid) {
                    float *dst = dW + (long)mrow * ldw + n0 + c0;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + c0 + i < n) atomicAdd(dst + i, v[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
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



inline int make_map(CUtensorMap *map, const float *base, long rows, int cols, long ld) {
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)kKRows};
    const cuuint32_t estr[2] = {1, 1};
    return encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace gwt

// PASNL_WGRAD_TMA=0 disables the kernel (A/B measurements).
extern "C" int pasnl_gemm_wgrad_tma_eligible(long rows, int m, int n, const void *X, long ldx, const void *Y, long ldy) {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_WGRAD_TMA");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (!on || gwt::encode_fn() == nullptr) return 0;
    if ((ldx % 4) != 0 || (ldy % 4) != 0 || (reinterpret_cast<uintptr_t>(X) & 15u) != 0 ||
        (reinterpret_cast<uintptr_t>(Y) & 15u) != 0)
        return 0;
    if (m < 32 || n < 32 || rows < 2048 || rows >= 2147483647L - 64) return 0;
    return 1;
}

// dW (m, n) += X (rows, m)^T . Y (rows, n); dw_zeroed == 0: the library clears dW first.
extern "C" int pasnl_gemm_wgrad_tma(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW,
                                    long ldw, int dw_zeroed, const float *x_scale, const float *x_shift, int x_act,
                                    void *stream) {
    PASNL_CHECK_ARG((x_scale == nullptr) == (x_shift == nullptr) && x_act >= 0 && x_act <= 2);
    PASNL_CHECK_ARG(rows >= 1 && m >= 1 && n >= 1 && ldx >= m && ldy >= n && ldw >= n && X && Y && dW);
    PASNL_CHECK_ARG(pasnl_gemm_wgrad_tma_eligible(rows, m, n, X, ldx, Y, ldy));
    cudaStream_t st = (cudaStream_t)stream;
    if (!dw_zeroed) {
        cudaError_t me = cudaMemset2DAsync(dW, sizeof(float) * (size_t)ldw, 0, sizeof(float) * (size_t)n, (size_t)m, st);
        if (me != cudaSuccess) return (int)me;
    }
    CUtensorMap xmap, ymap;
    if (gwt::make_map(&xmap, X, rows, m, ldx) || gwt::make_map(&ymap, Y, rows, n, ldy)) return PASNL_EINVAL;
    const int nq = n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 96 ? 96 : 128));
    const int mt = (m + 127) / 128, ntl = (n + nq - 1) / nq;
    long want = gwt::sm_count() / ((long)mt * ntl);
    if (want < 1) want = 1;
    long slab = (rows + want - 1) / want;
    if (slab < 512) slab = 512;
    slab = (slab + gwt::kKRows - 1) / gwt::kKRows * gwt::kKRows;
    const long slabs = (rows + slab - 1) / slab;
    const dim3 grid((unsigned)slabs, (unsigned)mt, (unsigned)ntl);
    const size_t smem = (size_t)(gwt::kRaw + gwt::kLo) * gwt::kRawStage + 1024;
    cudaError_t e = cudaFuncSetAttribute(gwt::gemm_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gwt::gemm_wgrad_tma_kernel<<<grid, gwt::kThreads, smem, st>>>(xmap, ymap, rows, slab, m, n, nq, dW, ldw, x_scale, x_shift,
                                                                  x_act);
    return pasnl_launch_status();
}
