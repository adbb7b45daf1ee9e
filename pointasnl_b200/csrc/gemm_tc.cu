// gemm_tc.cu -- fp32-accurate "rows" GEMM on tcgen05 tensor cores (3xTF32 split, TMEM accumulators), sm_100a.
//
//     C[rows, n] = A[rows, k] . Bt[n, k]^T          (A and Bt row-major: the contraction index is contiguous in both)
//
// This is the shape of every 1x1-conv / fully-connected layer around the point-set ops:
// forward y = x.W (Bt = W^T), input gradient dx = dy.W^T (Bt = W).  rows = B*S*K is 10^4..10^6, k and n are 3..4096.
// cuBLAS runs these as fp32 SIMT SGEMMs (~50 TFLOP/s, 5.5 ms of a 11.5 ms cls step); a single TF32 pass would break the
// 1e-4 parity bar, so each product is split a = a_hi + a_lo (a_hi = a truncated to TF32) and accumulated as
// a_hi.b_hi + a_hi.b_lo + a_lo.b_hi in fp32 TMEM (error ~2^-21 relative).  At 3 MMAs per k-step the tensor pipe still
// outruns HBM for these skinny shapes, so the kernel is memory bound: bytes = 4*rows*(k+n).
//
// CTA = 256 threads = one 128-row x NT-column output tile; k is consumed in 16-wide chunks through TWO operand stages:
//   all threads load + split the A (128 x 16) and Bt (NT x 16) chunks into the canonical no-swizzle K-major UMMA
//   layout [row/8][k/4][row%8][4] (hi and lo copies); one thread issues 2 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=NT) and a
//   tcgen05.commit to the stage's mbarrier; the MMAs run asynchronously while the CTA loads the next chunk into the
//   other stage (a stage is rewritten only after its commit has arrived).  After the last chunk the accumulator rows
//   leave TMEM through a shared-memory panel so that global stores are coalesced.
#include "common.cuh"

using namespace pasnl;

namespace gtc {

constexpr int kM = 128;
constexpr int kKC = 16;     // k chunk per stage: small stages -> 3-4 co-resident CTAs per SM keep ~48 KB of loads in flight

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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

// 4 consecutive k of one row, zero past the matrix edge; vector load when the row pitch allows it
template <bool VEC4>
__device__ __forceinline__ float4 load4(const float *__restrict__ base, long row, long rows, long ld, int k0, int k) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) {
        const float *p = base + row * ld + k0;
        if constexpr (VEC4) {
            if (k0 + 3 < k) return __ldg(reinterpret_cast<const float4 *>(p));
        }
        if (k0 < k) x.x = __ldg(p);
        if (k0 + 1 < k) x.y = __ldg(p + 1);
        if (k0 + 2 < k) x.z = __ldg(p + 2);
        if (k0 + 3 < k) x.w = __ldg(p + 3);
    }
    return x;
}
__device__ __forceinline__ void store_split4(float *hi, float *lo, int rows, int row, int kc, float4 x) {
    // hi = x with the 13 low mantissa bits cleared (exactly a TF32 value), lo = x - hi (exact in fp32, |lo| < 2^-10 |x|).
    // One LOP + one FADD per element: the F2F conversions (cvt.rna.tf32) run on the 16-lane XU pipe and were the
    // bottleneck of the first revision.  The tensor core drops lo's bits below TF32: 2^-20 relative to x.
    float4 h, l;
    h.x = tf32_hi(x.x); h.y = tf32_hi(x.y); h.z = tf32_hi(x.z); h.w = tf32_hi(x.w);
    l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
    const int o = (kc * rows + row) * 4;
    *reinterpret_cast<float4 *>(hi + o) = h;
    *reinterpret_cast<float4 *>(lo + o) = l;
}

// Operand tile layout (canonical no-swizzle K-major, row-group major):  [row/8][kc = k/4][row%8][4 floats]
//   core matrix = 8 rows x 16 bytes (128 contiguous bytes); LBO (next core matrix along k) = 128 bytes;
//   SBO (next 8-row group) = (kKC/4) * 128 bytes.
// A warp fills 8 rows x 4 k-chunks per instruction: lane = (row%8) + 8*(kc%4).  Global side: 8 rows x 64 contiguous
// bytes (8 L1 tags, 16 sectors) instead of 32 different rows; shared side: 4 x 128 contiguous bytes, conflict free.
constexpr int kSBO = (kKC / 4) * 128;   // bytes
__device__ __forceinline__ int tile_off(int row, int kc) { return ((row >> 3) * (kKC / 4) + kc) * 32 + (row & 7) * 4; }

template <int ROWS>
struct TileRegs {
    static constexpr int ITEMS = (ROWS / 8) * (kKC / 16);   // (row group, 16-wide k slab) pairs, 32 float4 each
    static constexpr int PER_WARP = (ITEMS + 7) / 8;
    float4 v[PER_WARP];
};

// global -> registers (issued one chunk ahead of its use: the loads fly while the previous chunk is split and MMA'd)
template <int ROWS, bool VEC4>
__device__ __forceinline__ void tile_fetch(TileRegs<ROWS> &tr, const float *__restrict__ src, long row_base,
                                           long row_limit, long ld, int kb, int k, int warp, int lane) {
    const int r8 = lane & 7, kq = lane >> 3;
#pragma unroll
    for (int j = 0; j < TileRegs<ROWS>::PER_WARP; ++j) {
        const int item = warp + 8 * j;
        const int rg = item % (ROWS / 8), h = item / (ROWS / 8);
        tr.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < TileRegs<ROWS>::ITEMS)
            tr.v[j] = load4<VEC4>(src, row_base + rg * 8 + r8, row_limit, ld, kb + (kq + 4 * h) * 4, k);
    }
}
// registers -> split -> shared memory operand tiles.  hi = a & mask, lo = a - hi: one LOP + one FADD per element.
template <int ROWS>
__device__ __forceinline__ void tile_store(const TileRegs<ROWS> &tr, float *hi, float *lo, int warp, int lane) {
    const int r8 = lane & 7, kq = lane >> 3;
#pragma unroll
    for (int j = 0; j < TileRegs<ROWS>::PER_WARP; ++j) {
        const int item = warp + 8 * j;
        const int rg = item % (ROWS / 8), h = item / (ROWS / 8);
        if (item < TileRegs<ROWS>::ITEMS) {
            const float4 x = tr.v[j];
            float4 hh, ll;
            hh.x = tf32_hi(x.x); hh.y = tf32_hi(x.y); hh.z = tf32_hi(x.z); hh.w = tf32_hi(x.w);
            ll.x = x.x - hh.x; ll.y = x.y - hh.y; ll.z = x.z - hh.z; ll.w = x.w - hh.w;
            const int o = tile_off(rg * 8 + r8, kq + 4 * h);
            *reinterpret_cast<float4 *>(hi + o) = hh;
            *reinterpret_cast<float4 *>(lo + o) = ll;
        }
    }
}

constexpr int tmem_cols(int nt) { return nt < 32 ? 32 : nt; }   // power of two >= 32 (NT is 32/64/128/256)

constexpr int kThreads = 256;   // 8 warps: all load/split; warp w reads TMEM lanes 32*(w%4).. in the epilogue

template <int NT, bool VEC4>
__global__ void __launch_bounds__(kThreads) gemm_rows_kernel(long rows, int k, int n, const float *__restrict__ A,
                                                             long lda, const float *__restrict__ Bt, long ldb,
                                                             float *__restrict__ C, long ldc,
                                                             double *__restrict__ stats) {
    constexpr int kStage = 2 * kM * kKC + 2 * NT * kKC;            // floats per operand stage (A hi/lo, B hi/lo)
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar[2];                                    // "stage s has been consumed by the tensor core"
    __shared__ uint32_t tmem_slot;

    const int t = threadIdx.x, warp = t >> 5;
    const long row0 = (long)blockIdx.x * kM;
    const int n0 = blockIdx.y * NT;
    const int nvalid = min(NT, n - n0);

    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(tmem_cols(NT))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = make_idesc(kM, NT);

    // Software pipeline over 16-wide k-chunks with two operand stages: while the tensor core consumes stage s
    // (asynchronously -- nobody waits for the MMAs of chunk i before loading chunk i+1), all 256 threads load and
    // split the next chunk into stage s^1.  A stage is only rewritten after the commit of the MMAs that read it.
    const int nchunks = (k + kKC - 1) / kKC;
    const int lane = t & 31;
    uint32_t ph0 = 0, ph1 = 0;
    // register prefetch, distance 2: chunk i+2 is requested while chunk i is split/stored, so two chunks' worth of
    // global loads per thread are always in flight (the loop is otherwise one memory latency per chunk)
    TileRegs<kM> ra[2];
    TileRegs<NT> rb[2];
    tile_fetch<kM, VEC4>(ra[0], A, row0, rows, lda, 0, k, warp, lane);
    tile_fetch<NT, VEC4>(rb[0], Bt, n0, (long)n0 + nvalid, ldb, 0, k, warp, lane);
    if (nchunks > 1) {
        tile_fetch<kM, VEC4>(ra[1], A, row0, rows, lda, kKC, k, warp, lane);
        tile_fetch<NT, VEC4>(rb[1], Bt, n0, (long)n0 + nvalid, ldb, kKC, k, warp, lane);
    }
    auto step = [&](int i, TileRegs<kM> &rga, TileRegs<NT> &rgb) {
        const int sidx = i & 1;
        float *Ahi = smem + sidx * kStage, *Alo = Ahi + kM * kKC;
        float *Bhi = Alo + kM * kKC, *Blo = Bhi + NT * kKC;
        if (i >= 2) {
            if (sidx == 0) { mbar_wait(&bar[0], ph0); ph0 ^= 1u; }
            else           { mbar_wait(&bar[1], ph1); ph1 ^= 1u; }
        }
        tile_store<kM>(rga, Ahi, Alo, warp, lane);
        tile_store<NT>(rgb, Bhi, Blo, warp, lane);
        if (i + 2 < nchunks) {
            tile_fetch<kM, VEC4>(rga, A, row0, rows, lda, (i + 2) * kKC, k, warp, lane);
            tile_fetch<NT, VEC4>(rgb, Bt, n0, (long)n0 + nvalid, ldb, (i + 2) * kKC, k, warp, lane);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks) {
                const uint64_t ah = make_desc(smem_u32(Ahi) + ks * 256, 128, kSBO);
                const uint64_t al = make_desc(smem_u32(Alo) + ks * 256, 128, kSBO);
                const uint64_t bh = make_desc(smem_u32(Bhi) + ks * 256, 128, kSBO);
                const uint64_t bl = make_desc(smem_u32(Blo) + ks * 256, 128, kSBO);
                mma_tf32(tmem, ah, bh, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                mma_tf32(tmem, ah, bl, idesc, 1);
                mma_tf32(tmem, al, bh, idesc, 1);
            }
            mma_commit(&bar[sidx]);
        }
    };
    for (int i = 0; i < nchunks; i += 2) {
        step(i, ra[0], rb[0]);
        if (i + 1 < nchunks) step(i + 1, ra[1], rb[1]);
    }
    // the last commit completes after every earlier MMA of this thread: one wait covers the whole accumulator
    if (((nchunks - 1) & 1) == 0) mbar_wait(&bar[0], ph0);
    else mbar_wait(&bar[1], ph1);
    tc_fence_after();

    // ---- epilogue: TMEM rows -> shared-memory panel (<= 64 columns at a time) -> coalesced global stores.
    // warps w and w+4 share TMEM lanes 32*(w%4)..; with a 64-column panel each takes one 32-column half.
    constexpr int PANEL = NT < 64 ? NT : 64, PP = PANEL + 4;
    float *stage = smem;                                          // 128 x (PANEL+4) floats, inside operand stage 0
    const int lrow = (warp & 3) * 32 + (t & 31);                  // the accumulator row this thread can read
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    for (int p0 = 0; p0 < NT; p0 += PANEL) {
        const int c0 = (warp >> 2) * 32;
        if (c0 < PANEL) {
            float v[32];
            tmem_ld32(tmem + lane_sel + p0 + c0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4 *>(stage + lrow * PP + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        __syncthreads();
        if (stats != nullptr && t < PANEL && n0 + p0 + t < n) {
            // fused batch-norm statistics: per-column sum and sum of squares of this tile's valid rows, accumulated
            // around a per-tile shift (row 0) in fp32 and converted to plain sums in fp64 before the atomics
            const int nv = (int)min((long)kM, rows - row0);
            const float shift = stage[t];
            float s1 = 0.f, s2 = 0.f;
            for (int r = 0; r < nv; ++r) {
                const float dlt = stage[r * PP + t] - shift;
                s1 += dlt;
                s2 = fmaf(dlt, dlt, s2);
            }
            const double sh = (double)shift, d1 = (double)s1;
            atomicAdd(stats + n0 + p0 + t, d1 + nv * sh);
            atomicAdd(stats + n + n0 + p0 + t, (double)s2 + 2.0 * sh * d1 + nv * sh * sh);
        }
        constexpr int CPR = PANEL / 4;                            // float4 chunks per row
        for (int e = t; e < kM * CPR; e += kThreads) {
            const int r = e / CPR, c = (e - r * CPR) * 4;
            const long grow = row0 + r;
            const int gcol = n0 + p0 + c;
            if (grow < rows && gcol < n) {
                const float4 v = *reinterpret_cast<const float4 *>(stage + r * PP + c);
                float *dst = C + grow * ldc + gcol;
                if (gcol + 3 < n && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
                    *reinterpret_cast<float4 *>(dst) = v;
                } else {
                    dst[0] = v.x;
                    if (gcol + 1 < n) dst[1] = v.y;
                    if (gcol + 2 < n) dst[2] = v.z;
                    if (gcol + 3 < n) dst[3] = v.w;
                }
            }
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols(NT)) : "memory");
}

template <int NT>
int launch(long rows, int k, int n, const float *A, long lda, const float *Bt, long ldb, float *C, long ldc,
           double *stats, cudaStream_t st) {
    const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(Bt) & 15u) == 0);
    const size_t smem = sizeof(float) * 2 * (size_t)(2 * kM * kKC + 2 * NT * kKC);   // two operand stages
    dim3 grid((unsigned)((rows + kM - 1) / kM), (unsigned)((n + NT - 1) / NT));
    cudaError_t e;
    if (vec) {
        auto kern = gemm_rows_kernel<NT, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kThreads, smem, st>>>(rows, k, n, A, lda, Bt, ldb, C, ldc, stats);
    } else {
        auto kern = gemm_rows_kernel<NT, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kThreads, smem, st>>>(rows, k, n, A, lda, Bt, ldb, C, ldc, stats);
    }
    return pasnl_launch_status();
}

// ================================================================ weight-gradient GEMM:  dW[m, n] = sum_r X[r, m] * Y[r, n]
// Both operands arrive row-major with the contraction index r as the SLOW one.  (Feeding them as MN-major UMMA operands
// with SWIZZLE_NONE produced zeros on hardware -- 32-bit MN-major apparently wants the 128B_BASE32B swizzle -- so the
// tiles are transposed while they are written to shared memory and the proven K-major descriptors are reused.)
// One tcgen05.mma.kind::tf32 consumes 8 r-rows.  The row range is split into slabs over CTAs; each CTA accumulates its
// 128 x NT partial in TMEM and adds it into dW (pre-zeroed) with vector reductions.
constexpr int kRC = 16;     // r rows per stage

template <int MN>
struct WTileRegs {
    static constexpr int ITEMS = (kRC / 8) * (MN / 16);   // (8-row block of r, group of 16 mn) pairs, 32 float4 each
    static constexpr int PER_WARP = (ITEMS + 7) / 8;
    float4 v[PER_WARP];
};
template <int MN, bool VEC4>
__device__ __forceinline__ void wtile_fetch(WTileRegs<MN> &tr, const float *__restrict__ src, long r0, long r_end,
                                            long ld, int c0, int cols, int warp, int lane) {
    const int r8 = lane & 7, cq = lane >> 3;
#pragma unroll
    for (int j = 0; j < WTileRegs<MN>::PER_WARP; ++j) {
        const int item = warp + 8 * j;
        const int rb = item % (kRC / 8), cg = item / (kRC / 8);
        tr.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < WTileRegs<MN>::ITEMS)   // load4(base, row, row_limit, ld, col0, col_limit): 4 consecutive columns
            tr.v[j] = load4<VEC4>(src, r0 + rb * 8 + r8, r_end, ld, c0 + (cg * 4 + cq) * 4, cols);
    }
}
// registers -> 4x4 transpose across 4 lanes -> split -> store.  The operand the tensor core sees is K-major with
// rows = m (or n) and k = r: element (mn, r) at tile_off(mn, r/4) + r%4 (same proven descriptors as the rows GEMM).
// A lane arrives with 4 consecutive mn at ONE r; the 4 lanes that hold r = 4a..4a+3 of the same mn group swap values
// with two shfl.xor rounds so that each ends up with 4 consecutive r of ONE mn -- exactly one 16-byte core-matrix row,
// written with a conflict-free STS.128 (a warp covers 4 x 128 contiguous bytes).
template <int MN>
__device__ __forceinline__ void wtile_store(const WTileRegs<MN> &tr, float *hi, float *lo, int warp, int lane) {
    const int r8 = lane & 7, cq = lane >> 3;
    const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
    for (int j = 0; j < WTileRegs<MN>::PER_WARP; ++j) {
        const int item = warp + 8 * j;
        const int rb = item % (kRC / 8), cg = item / (kRC / 8);
        float v0 = tr.v[j].x, v1 = tr.v[j].y, v2 = tr.v[j].z, v3 = tr.v[j].w;   // [mn0 + e] at r = rb*8 + r8
        {   // 2x2 blocks on lane bit 0 / element bit 0
            const float s0 = b0 ? v0 : v1, s1 = b0 ? v2 : v3;
            const float g0 = __shfl_xor_sync(FULL_MASK, s0, 1), g1 = __shfl_xor_sync(FULL_MASK, s1, 1);
            if (b0) { v0 = g0; v2 = g1; } else { v1 = g0; v3 = g1; }
        }
        {   // 2x2 blocks on lane bit 1 / element bit 1
            const float s0 = b1 ? v0 : v2, s1 = b1 ? v1 : v3;
            const float g0 = __shfl_xor_sync(FULL_MASK, s0, 2), g1 = __shfl_xor_sync(FULL_MASK, s1, 2);
            if (b1) { v0 = g0; v1 = g1; } else { v2 = g0; v3 = g1; }
        }
        // now (v0..v3) = element mn0 + (lane&3) at r = rb*8 + 4*((lane>>2)&1) + 0..3
        if (item < WTileRegs<MN>::ITEMS) {
            float4 hh, ll;
            hh.x = tf32_hi(v0); hh.y = tf32_hi(v1); hh.z = tf32_hi(v2); hh.w = tf32_hi(v3);
            ll.x = v0 - hh.x; ll.y = v1 - hh.y; ll.z = v2 - hh.z; ll.w = v3 - hh.w;
            const int mn = (cg * 4 + cq) * 4 + (lane & 3);
            const int o = tile_off(mn, rb * 2 + ((r8 >> 2) & 1));
            *reinterpret_cast<float4 *>(hi + o) = hh;
            *reinterpret_cast<float4 *>(lo + o) = ll;
        }
    }
}

template <int NT, bool VEC4>
__global__ void __launch_bounds__(kThreads) gemm_wgrad_kernel(long rows, int m, int n, long slab,
                                                              const float *__restrict__ X, long ldx,
                                                              const float *__restrict__ Y, long ldy,
                                                              float *__restrict__ dW, long ldw) {
    constexpr int kStage = 2 * kRC * kM + 2 * kRC * NT;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m0 = blockIdx.y * kM, n0 = blockIdx.z * NT;
    const long r_begin = (long)blockIdx.x * slab, r_end = min(rows, r_begin + slab);
    if (r_begin >= r_end) return;

    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(tmem_cols(NT))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = make_idesc(kM, NT);
    static_assert(kRC == kKC, "the transposed operand tiles reuse tile_off() of the rows GEMM");

    const int nchunks = (int)((r_end - r_begin + kRC - 1) / kRC);
    uint32_t ph0 = 0, ph1 = 0;
    WTileRegs<kM> ra;
    WTileRegs<NT> rb;
    wtile_fetch<kM, VEC4>(ra, X, r_begin, r_end, ldx, m0, m, warp, lane);
    wtile_fetch<NT, VEC4>(rb, Y, r_begin, r_end, ldy, n0, n, warp, lane);
    for (int i = 0; i < nchunks; ++i) {
        const int sidx = i & 1;
        float *Ahi = smem + sidx * kStage, *Alo = Ahi + kRC * kM;
        float *Bhi = Alo + kRC * kM, *Blo = Bhi + kRC * NT;
        if (i >= 2) {
            if (sidx == 0) { mbar_wait(&bar[0], ph0); ph0 ^= 1u; }
            else           { mbar_wait(&bar[1], ph1); ph1 ^= 1u; }
        }
        wtile_store<kM>(ra, Ahi, Alo, warp, lane);
        wtile_store<NT>(rb, Bhi, Blo, warp, lane);
        if (i + 1 < nchunks) {
            const long rn = r_begin + (long)(i + 1) * kRC;
            wtile_fetch<kM, VEC4>(ra, X, rn, r_end, ldx, m0, m, warp, lane);
            wtile_fetch<NT, VEC4>(rb, Y, rn, r_end, ldy, n0, n, warp, lane);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kRC / 8; ++ks) {
                const uint64_t ah = make_desc(smem_u32(Ahi) + ks * 256, 128, kSBO);
                const uint64_t al = make_desc(smem_u32(Alo) + ks * 256, 128, kSBO);
                const uint64_t bh = make_desc(smem_u32(Bhi) + ks * 256, 128, kSBO);
                const uint64_t bl = make_desc(smem_u32(Blo) + ks * 256, 128, kSBO);
                mma_tf32(tmem, ah, bh, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                mma_tf32(tmem, ah, bl, idesc, 1);
                mma_tf32(tmem, al, bh, idesc, 1);
            }
            mma_commit(&bar[sidx]);
        }
    }
    if (((nchunks - 1) & 1) == 0) mbar_wait(&bar[0], ph0);
    else mbar_wait(&bar[1], ph1);
    tc_fence_after();

    // ---- epilogue: this CTA's partial (row = m index) is added into dW; warps w and w+4 split the columns
    const int mrow = m0 + (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    for (int c0 = (warp >> 2) * 32; c0 < NT; c0 += 64) {
        float v[32];
        tmem_ld32(tmem + lane_sel + c0, v);
        if (mrow < m) {
            float *dst = dW + (long)mrow * ldw + n0 + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (n0 + c0 + i < n) atomicAdd(dst + i, v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols(NT)) : "memory");
}

template <int NT>
int launch_wgrad(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW, long ldw,
                 cudaStream_t st) {
    const bool vec = (ldx % 4 == 0) && (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(Y) & 15u) == 0);
    const size_t smem = sizeof(float) * 2 * (size_t)(2 * kRC * kM + 2 * kRC * NT);
    const int mt = (m + kM - 1) / kM, nt = (n + NT - 1) / NT;
    // row slabs: enough CTAs for ~3 per SM over all (m, n) tiles, at least 256 rows each
    long want = (148L * 3 + (long)mt * nt - 1) / ((long)mt * nt);
    if (want < 1) want = 1;
    long slab = (rows + want - 1) / want;
    if (slab < 256) slab = 256;
    slab = (slab + kRC - 1) / kRC * kRC;
    const long slabs = (rows + slab - 1) / slab;
    dim3 grid((unsigned)slabs, (unsigned)mt, (unsigned)nt);
    cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)m * ldw, st);
    if (e != cudaSuccess) return (int)e;
    if (vec) {
        auto kern = gemm_wgrad_kernel<NT, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kThreads, smem, st>>>(rows, m, n, slab, X, ldx, Y, ldy, dW, ldw);
    } else {
        auto kern = gemm_wgrad_kernel<NT, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kThreads, smem, st>>>(rows, m, n, slab, X, ldx, Y, ldy, dW, ldw);
    }
    return pasnl_launch_status();
}

}  // namespace gtc

extern "C" int pasnl_gemm_rows_tf32x3(long rows, int k, int n, const float *A, long lda, const float *Bt, long ldb,
                                      float *C, long ldc, double *col_stats, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && k >= 1 && n >= 1 && lda >= k && ldb >= k && ldc >= n);
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(A && Bt && C && rows / 128 < 2147483647L && n <= 256 * 65535);
    cudaStream_t st = (cudaStream_t)stream;
    // UMMA N must be a multiple of 16 (M = 128): the tile is padded in shared memory, stores are bounds-checked
    if (n <= 32) return gtc::launch<32>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, st);
    if (n <= 64) return gtc::launch<64>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, st);
    if (n <= 128) return gtc::launch<128>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, st);
    return gtc::launch<256>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, st);
}

extern "C" int pasnl_gemm_wgrad_tf32x3(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy,
                                       float *dW, long ldw, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && m >= 1 && n >= 1 && ldx >= m && ldy >= n && ldw >= n);
    PASNL_CHECK_ARG(dW && (rows == 0 || (X && Y)) && ldw == n && m <= 128 * 65535 && n <= 256 * 65535);
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)m * ldw, st);
        return (int)e;
    }
    if (n <= 32) return gtc::launch_wgrad<32>(rows, m, n, X, ldx, Y, ldy, dW, ldw, st);
    if (n <= 64) return gtc::launch_wgrad<64>(rows, m, n, X, ldx, Y, ldy, dW, ldw, st);
    return gtc::launch_wgrad<128>(rows, m, n, X, ldx, Y, ldy, dW, ldw, st);
}
