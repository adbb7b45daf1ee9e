// gemm_tc.cu -- the fp32-accurate GEMMs of the 1x1-conv / fully-connected layers on tcgen05 tensor cores (3xTF32 split,
// TMEM accumulators), sm_100a.  Two products:
//
//     rows GEMM     C[rows, n]  = A[rows, k] . Bt[n, k]^T     forward y = x.W (Bt = W^T), input gradient dx = dy.W^T
//     weight grad   dW[m, n]    = sum_r X[r, m] * Y[r, n]
//
// rows = B*S*K is 10^4..10^6, k / n / m are 3..4096.  cuBLAS runs these as fp32 SIMT SGEMMs (~50 TFLOP/s, 5.5 ms of the
// first 11.5 ms cls step); a single TF32 pass would break the 1e-4 parity bar, so each product is split a = a_hi + a_lo
// (a_hi = a truncated to TF32) and accumulated as a_hi.b_hi + a_hi.b_lo + a_lo.b_hi in fp32 TMEM (~2^-21 relative).
//
// Kernels in this file, and when pasnl_gemm_*_tf32x3 picks them:
//   gemm_rows_persist_kernel<NT>   rows >= 2*128*SMs, k <= 144, n <= 160: one CTA per SM, weights resident in shared
//                                  memory, producer / MMA / epilogue warps over mbarriers, two TMEM accumulators
//   gemm_rows_kernel<NT>           everything else: one 128 x NT tile per CTA, two operand stages (below)
//   gemm_wgrad_tm_kernel           long reductions: one operand is fed through TMEM (tcgen05.st + TS-mode MMA)
//   gemm_wgrad_kernel<NT>          short reductions: both operands transposed into shared memory, ~3 CTAs per SM
//   wgrad_narrow_kernel<MP,WQ>     one operand with <= 16 columns: fp32 FMA streaming pass
// What bounds them (ncu, profiles/r01b_*): the L1 / shared-memory data pipe -- operand STS plus the tensor core's own
// shared-memory operand reads (3 MMAs per k-step re-read both operands) -- not HBM and not the tensor pipe.
//
// gemm_rows_kernel: CTA = 256 threads = one 128-row x NT-column output tile; k is consumed in 16-wide chunks through
//   TWO operand stages: all threads load + split the A (128 x 16) and Bt (NT x 16) chunks into the canonical no-swizzle
//   K-major UMMA layout [row/8][k/4][row%8][4] (hi and lo copies); one thread issues 2 k-steps x 3
//   tcgen05.mma.kind::tf32 (M=128, N=NT) and a tcgen05.commit to the stage's mbarrier; the MMAs run asynchronously
//   while the CTA loads the next chunk into the other stage (a stage is rewritten only after its commit has arrived).
//   After the last chunk the accumulator rows leave TMEM through a shared-memory panel so that global stores are
//   coalesced.
#include <cstdlib>

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
// one lane of a converged warp (ELECT): unlike `lane == 0` the compiler knows the guarded region runs on exactly one
// lane and issues the uniform-datapath tcgen05 instructions without a per-instruction election loop
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
// The same fetch with the address arithmetic hoisted out of the k loop: one pointer per item, computed once per tile.
// `fast` = every row of the tile exists and loads are 16-byte aligned; only a ragged last k chunk takes the checked path.
template <int ROWS>
struct TilePtrs {
    const float *p[TileRegs<ROWS>::PER_WARP];
    bool fast;
};
template <int ROWS>
__device__ __forceinline__ void tile_ptrs(TilePtrs<ROWS> &tp, const float *__restrict__ src, long row_base, long row_limit,
                                          long ld, bool vec4, int warp, int lane) {
    static_assert(kKC == 16, "one 16-wide k slab per item");
    const int r8 = lane & 7, kq = lane >> 3;
#pragma unroll
    for (int j = 0; j < TileRegs<ROWS>::PER_WARP; ++j) {
        const int item = warp + 8 * j;
        tp.p[j] = src + (row_base + (item % (ROWS / 8)) * 8 + r8) * ld + kq * 4;
    }
    tp.fast = vec4 && (row_base + ROWS <= row_limit);
}
template <int ROWS, bool VEC4>
__device__ __forceinline__ void tile_fetch_fast(TileRegs<ROWS> &tr, const TilePtrs<ROWS> &tp, const float *__restrict__ src,
                                                long row_base, long row_limit, long ld, int kb, int k, int warp, int lane) {
    if (tp.fast && kb + kKC <= k) {
#pragma unroll
        for (int j = 0; j < TileRegs<ROWS>::PER_WARP; ++j) {
            tr.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (warp + 8 * j < TileRegs<ROWS>::ITEMS) tr.v[j] = __ldg(reinterpret_cast<const float4 *>(tp.p[j] + kb));
        }
    } else {
        tile_fetch<ROWS, VEC4>(tr, src, row_base, row_limit, ld, kb, k, warp, lane);
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

constexpr int tmem_cols(int nt) {   // TMEM allocations are powers of two >= 32 columns
    int c = 32;
    while (c < nt) c *= 2;
    return c;
}

constexpr int kThreads = 256;   // 8 warps: all load/split; warp w reads TMEM lanes 32*(w%4).. in the epilogue

template <int NT, bool VEC4>
__global__ void __launch_bounds__(kThreads) gemm_rows_kernel(long rows, int k, int n, const float *__restrict__ A,
                                                             long lda, const float *__restrict__ Bt, long ldb,
                                                             float *__restrict__ C, long ldc,
                                                             double *__restrict__ stats, int kslab) {
    constexpr int kStage = 2 * kM * kKC + 2 * NT * kKC;            // floats per operand stage (A hi/lo, B hi/lo)
    // split-k (gridDim.z > 1): this CTA contracts columns [z*kslab, (z+1)*kslab) and ADDS its tile into a zeroed C.
    // Long contractions are cut anyway (the TMEM adder truncates, ~3e-8*k relative: <= 1024-wide slabs summed with
    // round-to-nearest adds); running the slabs side by side fills the GPU when there are few row tiles.
    const bool split = gridDim.z > 1;
    {
        const int kz = blockIdx.z * kslab;
        A += kz;
        Bt += kz;
        k = min(kslab, k - kz);
    }
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
    TilePtrs<kM> pa;
    TilePtrs<NT> pb;
    tile_ptrs<kM>(pa, A, row0, rows, lda, VEC4, warp, lane);
    tile_ptrs<NT>(pb, Bt, n0, (long)n0 + nvalid, ldb, VEC4, warp, lane);
    tile_fetch_fast<kM, VEC4>(ra[0], pa, A, row0, rows, lda, 0, k, warp, lane);
    tile_fetch_fast<NT, VEC4>(rb[0], pb, Bt, n0, (long)n0 + nvalid, ldb, 0, k, warp, lane);
    if (nchunks > 1) {
        tile_fetch_fast<kM, VEC4>(ra[1], pa, A, row0, rows, lda, kKC, k, warp, lane);
        tile_fetch_fast<NT, VEC4>(rb[1], pb, Bt, n0, (long)n0 + nvalid, ldb, kKC, k, warp, lane);
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
            tile_fetch_fast<kM, VEC4>(rga, pa, A, row0, rows, lda, (i + 2) * kKC, k, warp, lane);
            tile_fetch_fast<NT, VEC4>(rgb, pb, Bt, n0, (long)n0 + nvalid, ldb, (i + 2) * kKC, k, warp, lane);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) {
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
                if (split) {
                    atomicAdd(dst, v.x);
                    if (gcol + 1 < n) atomicAdd(dst + 1, v.y);
                    if (gcol + 2 < n) atomicAdd(dst + 2, v.z);
                    if (gcol + 3 < n) atomicAdd(dst + 3, v.w);
                } else if (gcol + 3 < n && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
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
           double *stats, int kslab, cudaStream_t st) {
    const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(Bt) & 15u) == 0);
    const size_t smem = sizeof(float) * 2 * (size_t)(2 * kM * kKC + 2 * NT * kKC);   // two operand stages
    dim3 grid((unsigned)((rows + kM - 1) / kM), (unsigned)((n + NT - 1) / NT), (unsigned)((k + kslab - 1) / kslab));
    cudaError_t e;
    if (vec) {
        auto kern = gemm_rows_kernel<NT, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kThreads, smem, st>>>(rows, k, n, A, lda, Bt, ldb, C, ldc, stats, kslab);
    } else {
        auto kern = gemm_rows_kernel<NT, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kThreads, smem, st>>>(rows, k, n, A, lda, Bt, ldb, C, ldc, stats, kslab);
    }
    return pasnl_launch_status();
}

// ================================================================ persistent, warp-specialised rows GEMM (weights resident)
// For the layers that dominate a step (rows = B*S*K ~ 10^5..10^6, k <= 144, n <= 128) the weight matrix is tiny: its
// hi/lo operand tiles (<= 147 KB) are built ONCE per CTA and stay in shared memory while the CTA walks over row tiles.
// One CTA per SM, 13 warps:
//   warps 5..12  producers: stream A (128 x 16 chunks) global -> registers (6 chunks in flight per thread) -> split ->
//                 operand stage ring in shared memory, arrive on full[s]
//   warp  4      MMA issuer: waits full[s], issues 6 tcgen05.mma (2 k-steps x hi.hi, hi.lo, lo.hi) per chunk into one
//                 of TWO accumulators in TMEM, tcgen05.commit -> empty[s]; after the tile's last chunk commit -> tfull[b]
//   warps 0..3   epilogue: wait tfull[b], tcgen05.ld their 32 accumulator rows, release the accumulator (tempty[b]) and
//                 write C through a private 32x32 shared-memory panel (coalesced 128-byte rows), folding the batch-norm
//                 column statistics into per-lane fp64 registers that reach global memory once per CTA.
// The tile loop of the three roles only meets through mbarriers, so loads, tensor-core work and stores of neighbouring
// tiles overlap; the one-tile-per-CTA kernel above pays load latency, MMA latency and the store drain in sequence.
constexpr int kPThreads = 13 * 32;
constexpr int kPProducerWarps = 8;
constexpr int kPrefetch = 6;       // A chunks in flight per producer thread (6 x 8 KB per CTA ~ HBM latency x 26 B/clk)
constexpr int kMaxStages = 8;
constexpr int kPanelPitch = 36;    // floats: 32 columns + 4 (conflict-free 16-byte row writes)

__device__ __forceinline__ int btile_off(int row, int kc, int kq4) { return ((row >> 3) * kq4 + kc) * 32 + (row & 7) * 4; }

template <int NT>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_rows_persist_kernel(long rows, int k, int kpad, int n, int n0, int nvalid, long ntiles, int nstages,
                         const float *__restrict__ A, long lda, const float *__restrict__ Bt, long ldb,
                         float *__restrict__ C, long ldc, double *__restrict__ stats, int vec4) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int kq4 = kpad >> 2;                                   // 16-byte k groups per operand row
    float *Bhi = smem, *Blo = Bhi + NT * kpad;
    float *stages = Blo + NT * kpad;                             // nstages x [A hi 128x16 | A lo 128x16]
    float *panels = stages + (size_t)nstages * (2 * kM * kKC);   // 4 x 32 x kPanelPitch

    if (t == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], kPProducerWarps); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(tmem_cols(2 * NT))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // resident weights: Bt rows n0..n0+nvalid, k zero-padded to kpad, split once
        const int r8 = lane & 7, kq = lane >> 3;
        const int items = (NT / 8) * (kpad / 16);
        for (int item = warp; item < items; item += kPThreads / 32) {
            const int rg = item % (NT / 8), h = item / (NT / 8);
            const int row = rg * 8 + r8, kc = kq + 4 * h;
            const float4 x = vec4 ? load4<true>(Bt, (long)n0 + row, (long)n0 + nvalid, ldb, kc * 4, k)
                                  : load4<false>(Bt, (long)n0 + row, (long)n0 + nvalid, ldb, kc * 4, k);
            float4 hh, ll;
            hh.x = tf32_hi(x.x); hh.y = tf32_hi(x.y); hh.z = tf32_hi(x.z); hh.w = tf32_hi(x.w);
            ll.x = x.x - hh.x; ll.y = x.y - hh.y; ll.z = x.z - hh.z; ll.w = x.w - hh.w;
            const int o = btile_off(row, kc, kq4);
            *reinterpret_cast<float4 *>(Bhi + o) = hh;
            *reinterpret_cast<float4 *>(Blo + o) = ll;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int nchunks = kpad / kKC;
    const long my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

    if (warp >= 5) {
        // ------------------------------------------------------------------ producers
        // A warp's instruction stream is serial (~5 cycles per dependent instruction), so the per-chunk path is kept
        // to a few dozen instructions: every thread owns the same two (row, 16-byte k group) positions of each tile
        // and walks two pointers; index arithmetic and bounds checks only run on the ragged last tile / last chunk.
        const int pw = warp - 5;
        const int r8 = lane & 7, kq = lane >> 3;
        const int ra = pw * 8 + r8, rb = ra + 64;                 // this thread's two rows inside a tile
        const int oa = tile_off(ra, kq), ob = tile_off(rb, kq);   // and their float offsets inside an operand stage
        const int kfull = k / kKC;                                // chunks that need no k bounds check
        const long tstride = (long)gridDim.x * kM;
        long frow0 = (long)blockIdx.x * kM;                       // fetch cursor: tile row base, chunk, pointers
        int fc = 0;
        const float *pa = A + (frow0 + ra) * lda + kq * 4;
        const long hop = 64 * lda;
        bool tfullrows = frow0 + kM <= rows;
        long remaining_fetch = my_tiles * nchunks;
        float4 r[kPrefetch][2];
        auto fetch = [&](float4 (&dst)[2]) {
            if (vec4 && tfullrows && fc < kfull) {
                dst[0] = __ldg(reinterpret_cast<const float4 *>(pa));
                dst[1] = __ldg(reinterpret_cast<const float4 *>(pa + hop));
            } else if (vec4) {
                dst[0] = load4<true>(A, frow0 + ra, rows, lda, fc * kKC + kq * 4, k);
                dst[1] = load4<true>(A, frow0 + rb, rows, lda, fc * kKC + kq * 4, k);
            } else {
                dst[0] = load4<false>(A, frow0 + ra, rows, lda, fc * kKC + kq * 4, k);
                dst[1] = load4<false>(A, frow0 + rb, rows, lda, fc * kKC + kq * 4, k);
            }
            pa += kKC;
            if (++fc == nchunks) {
                fc = 0;
                frow0 += tstride;
                pa = A + (frow0 + ra) * lda + kq * 4;
                tfullrows = frow0 + kM <= rows;
            }
            --remaining_fetch;
        };
        auto split_store = [&](const float4 &x, float *hi, float *lo) {
            float4 hh, ll;
            hh.x = tf32_hi(x.x); hh.y = tf32_hi(x.y); hh.z = tf32_hi(x.z); hh.w = tf32_hi(x.w);
            ll.x = x.x - hh.x; ll.y = x.y - hh.y; ll.z = x.z - hh.z; ll.w = x.w - hh.w;
            *reinterpret_cast<float4 *>(hi) = hh;
            *reinterpret_cast<float4 *>(lo) = ll;
        };
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u)
            if (remaining_fetch > 0) fetch(r[u]);
        int s = 0;
        uint32_t par = 1;                                         // empty[] starts "already released"
        float *stage = stages;
        for (long left = my_tiles * nchunks; left > 0; left -= kPrefetch) {
#pragma unroll
            for (int u = 0; u < kPrefetch; ++u) {
                if (u < left) {
                    mbar_wait(&empty[s], par);
                    split_store(r[u][0], stage + oa, stage + kM * kKC + oa);
                    split_store(r[u][1], stage + ob, stage + kM * kKC + ob);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    if (remaining_fetch > 0) fetch(r[u]);
                    stage += 2 * kM * kKC;
                    if (++s == nstages) { s = 0; par ^= 1u; stage = stages; }
                }
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ MMA issuer (lane 0 issues; the warp stays converged)
        constexpr uint32_t idesc = make_idesc(kM, NT);
        const uint32_t sbo_b = (uint32_t)kq4 * 128u;
        int s = 0;
        uint32_t par = 0;
        for (long j = 0; j < my_tiles; ++j) {
            const int b = (int)(j & 1);
            mbar_wait(&tempty[b], (uint32_t)(((j >> 1) & 1) ^ 1));
            tc_fence_after();
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&full[s], par);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(stages + (size_t)s * (2 * kM * kKC)), a_lo = a_hi + kM * kKC * 4;
                    const uint32_t b_hi = smem_u32(Bhi) + (uint32_t)c * (kKC / 4) * 128u;
                    const uint32_t b_lo = smem_u32(Blo) + (uint32_t)c * (kKC / 4) * 128u;
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        const uint64_t ah = make_desc(a_hi + ks * 256, 128, kSBO);
                        const uint64_t al = make_desc(a_lo + ks * 256, 128, kSBO);
                        const uint64_t bh = make_desc(b_hi + ks * 256, 128, sbo_b);
                        const uint64_t bl = make_desc(b_lo + ks * 256, 128, sbo_b);
                        mma_tf32(tmem + b * NT, ah, bh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32(tmem + b * NT, ah, bl, idesc, 1);
                        mma_tf32(tmem + b * NT, al, bh, idesc, 1);
                    }
                    mma_commit(&empty[s]);
                    if (c == nchunks - 1) mma_commit(&tfull[b]);
                }
                __syncwarp();
                if (++s == nstages) { s = 0; par ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warp w owns accumulator rows 32w..32w+31)
        float *panel = panels + warp * 32 * kPanelPitch;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        double acc1[NT / 32], acc2[NT / 32];
#pragma unroll
        for (int q = 0; q < NT / 32; ++q) { acc1[q] = 0.0; acc2[q] = 0.0; }
        // whole launch: every column of the window is valid and rows of C can take 16-byte stores
        const bool fast_cols = (nvalid == NT) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C + n0) & 15u) == 0);
        const int pr = lane >> 3, pc = (lane & 7) * 4;            // coalesced read-back: 4 rows x 128 bytes per instruction
        const long step = 4 * ldc;
        for (long j = 0; j < my_tiles; ++j) {
            const int b = (int)(j & 1);
            const long row0 = ((long)blockIdx.x + j * gridDim.x) * kM + warp * 32;
            const int nv = (int)max(0L, min(32L, rows - row0));
            mbar_wait(&tfull[b], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            if (fast_cols && nv == 32) {
                // ---- full tile: no bounds logic, loads and stores batched (a serial LDS -> STG chain per row costs
                // ~100 cycles x 32 rows x 4 column groups and made this warp the pipeline's slowest stage)
                float *dst = C + (row0 + pr) * ldc + n0 + pc;
#pragma unroll
                for (int q = 0; q < NT / 32; ++q) {
                    float v[32];
                    tmem_ld32(tmem + lane_sel + b * NT + q * 32, v);
                    if (q == NT / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[b]);
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4 *>(panel + lane * kPanelPitch + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    __syncwarp();
                    if (stats != nullptr) {
                        // 32 rows in fp32 (4 independent chains), widened to fp64 before they meet other tiles
                        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) {
                            const float x = panel[rr * kPanelPitch + lane];
                            s1[rr & 3] += x;
                            s2[rr & 3] = fmaf(x, x, s2[rr & 3]);
                        }
                        acc1[q] += (double)((s1[0] + s1[1]) + (s1[2] + s1[3]));
                        acc2[q] += (double)((s2[0] + s2[1]) + (s2[2] + s2[3]));
                    }
                    float4 x[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        x[it] = *reinterpret_cast<const float4 *>(panel + (it * 4 + pr) * kPanelPitch + pc);
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        *reinterpret_cast<float4 *>(dst + it * step + q * 32) = x[it];
                    __syncwarp();
                }
                continue;
            }
#pragma unroll
            for (int q = 0; q < NT / 32; ++q) {
                float v[32];
                tmem_ld32(tmem + lane_sel + b * NT + q * 32, v);
                if (q == NT / 32 - 1) {                           // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[b]);
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4 *>(panel + lane * kPanelPitch + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
                const int col = q * 32 + lane;                    // column inside this launch's window
                if (stats != nullptr && nv > 0 && col < nvalid) {
                    const float shift = panel[lane];
                    float s1 = 0.f, s2 = 0.f;
                    for (int rr = 0; rr < nv; ++rr) {
                        const float dlt = panel[rr * kPanelPitch + lane] - shift;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                    }
                    const double sh = (double)shift, d1 = (double)s1;
                    acc1[q] += d1 + nv * sh;
                    acc2[q] += (double)s2 + 2.0 * sh * d1 + nv * sh * sh;
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3), c4 = (lane & 7) * 4;
                    const int gc = q * 32 + c4;
                    if (rr < nv && gc < nvalid) {
                        const float4 x = *reinterpret_cast<const float4 *>(panel + rr * kPanelPitch + c4);
                        float *dst = C + (row0 + rr) * ldc + n0 + gc;
                        if (gc + 3 < nvalid && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
                            *reinterpret_cast<float4 *>(dst) = x;
                        } else {
                            dst[0] = x.x;
                            if (gc + 1 < nvalid) dst[1] = x.y;
                            if (gc + 2 < nvalid) dst[2] = x.z;
                            if (gc + 3 < nvalid) dst[3] = x.w;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (stats != nullptr) {
#pragma unroll
            for (int q = 0; q < NT / 32; ++q) {
                const int col = q * 32 + lane;
                if (col < nvalid) {
                    atomicAdd(stats + n0 + col, acc1[q]);
                    atomicAdd(stats + n + n0 + col, acc2[q]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols(2 * NT)) : "memory");
}

inline bool persist_enabled() {            // PASNL_GEMM_PERSIST=0 forces the one-tile-per-CTA kernel (A/B measurements)
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_GEMM_PERSIST");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

inline bool wgrad_tm_enabled() {           // PASNL_WGRAD_TM=0: shared-memory-only weight-gradient kernel (A/B measurements)
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_WGRAD_TM");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
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

inline int min_tile_width() {               // PASNL_ROWS_MIN_NT: floor of the narrowing heuristic (default 32)
    static int v = 0;
    if (v == 0) {
        const char *e = getenv("PASNL_ROWS_MIN_NT");
        v = e ? atoi(e) : 32;
        if (v != 32 && v != 64 && v != 128 && v != 256) v = 32;
    }
    return v;
}

constexpr long kSmemBudget = 227 * 1024 - 2048;   // opt-in maximum minus the static barriers / slack

// operand stages that fit next to the resident weights and the panels (the kernel needs at least 2)
inline int persist_stages(int kpad, int nt) {
    const long fixed = 8L * nt * kpad + 4L * 4 * 32 * kPanelPitch;
    long st = (kSmemBudget - fixed) / (long)(sizeof(float) * 2 * kM * kKC);
    if (st > kMaxStages) st = kMaxStages;
    return (int)st;
}

template <int NT>
int launch_persist(long rows, int k, int n, int n0, int nvalid, const float *A, long lda, const float *Bt, long ldb,
                   float *C, long ldc, double *stats, cudaStream_t st) {
    const int kpad = (k + kKC - 1) / kKC * kKC;
    const int nstages = persist_stages(kpad, NT);
    const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(Bt) & 15u) == 0);
    const size_t smem = 8UL * NT * kpad + sizeof(float) * ((size_t)nstages * 2 * kM * kKC + 4 * 32 * kPanelPitch);
    const long ntiles = (rows + kM - 1) / kM;
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    auto kern = gemm_rows_persist_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kPThreads, smem, st>>>(rows, k, kpad, n, n0, nvalid, ntiles, nstages, A, lda, Bt, ldb, C, ldc, stats,
                                        vec ? 1 : 0);
    return pasnl_launch_status();
}

inline int launch_persist_cols(long rows, int k, int n, int n0, int nvalid, const float *A, long lda, const float *Bt,
                               long ldb, float *C, long ldc, double *stats, cudaStream_t st) {
    if (nvalid <= 32) return launch_persist<32>(rows, k, n, n0, nvalid, A, lda, Bt, ldb, C, ldc, stats, st);
    if (nvalid <= 64) return launch_persist<64>(rows, k, n, n0, nvalid, A, lda, Bt, ldb, C, ldc, stats, st);
    if (nvalid <= 128) return launch_persist<128>(rows, k, n, n0, nvalid, A, lda, Bt, ldb, C, ldc, stats, st);
    return launch_persist<160>(rows, k, n, n0, nvalid, A, lda, Bt, ldb, C, ldc, stats, st);
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
        if (warp == 0 && elect_one()) {
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

// ---------------------------------------------------------------- helpers of the long-reduction weight-gradient kernel
constexpr int kSubChunks = 64;     // 64 x 16 = 1024 rows per TMEM accumulation (the TMEM adder truncates: ~7e-9
                                   // relative per accumulated row, so partial sums are kept short)

__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    // for warps that wait most of the kernel: poll with back-off instead of spinning on the barrier
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(500);
    }
}

// ---------------------------------------------------------------- weight gradient with one operand fed through TMEM
// The L1/shared-memory data pipe is what bounds the kernels above (~0.7 wavefronts per clock per SM for STS + the
// tensor core's own operand reads).  For dW = X^T.Y both operands arrive [row][column] with the columns contiguous, so
// a warp whose lanes are 32 consecutive columns of one operand ("P") loads coalesced 128-byte lines AND already holds
// the layout tcgen05.st wants (TMEM lane = operand row = column of P, TMEM column = k = row index): P needs no
// transpose, no shared memory and no tensor-core shared-memory reads at all.  Only the other operand ("Q") is
// transposed into K-major shared-memory tiles.  D[p][q] = sum_r P[r][p0+p] * Q[r][q0+q]:
//   warps 0..3   epilogue (TMEM quadrant w): add finished 1024-row sub-slabs into dW
//   warps 4..11  P producers (quadrant (w-4)%4, rows 8*((w-4)/4).. of each chunk): 8 coalesced LDG.32 -> hi/lo ->
//                2 x tcgen05.st.32x32b.x8
//   warp  12     MMA issue: tcgen05.mma (A from TMEM, B descriptor), 2 k-steps x (hi.hi, hi.lo, lo.hi) per 16-row chunk
//   warps 13..20 Q producers: 4 rows x 128-byte loads, 4x4 lane transpose, split, STS.128 into the stage ring
constexpr int kTThreads = 21 * 32;
constexpr int kTQWarps = 8, kTPWarps = 8;
constexpr int kTMmaWarp = 4 + kTPWarps, kTQWarp0 = kTMmaWarp + 1;
constexpr int kTMaxNQ = 160;                 // two accumulators of <= 160 columns + 6 x 32 ring columns = 512
constexpr int kTRingCol = 2 * kTMaxNQ;
constexpr int kTMaxStages = 6;
constexpr int kTQItems = 3;                  // (4 rows x 32 columns) pieces of Q per producer warp and chunk (5 blocks max)
constexpr int kTPPrefetch = 4, kTQPrefetch = 2;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7]))
        : "memory");
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

__global__ void __launch_bounds__(kTThreads, 1)
gemm_wgrad_tm_kernel(long rows, long slab, int nstages, const float *__restrict__ P, long ldp, int pcols,
                     const float *__restrict__ Q, long ldq, int qcols, int nq, float *__restrict__ dW, long ldw,
                     int transposed, int vec4) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t full[kTMaxStages], empty[kTMaxStages], tfull[2], tempty[2];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int p0 = blockIdx.y * kM, q0 = blockIdx.z * nq;
    const long r_begin = (long)blockIdx.x * slab, r_end = min(rows, r_begin + slab);
    if (r_begin >= r_end) return;
    const int stage_floats = 2 * kRC * nq;                        // Q hi | Q lo

    if (t == 0) {
        for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], kTQWarps + kTPWarps); mbar_init(&empty[s], 1); }
        mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
        fence_mbar_init();
    }
    if (warp == kTMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int nchunks = (int)((r_end - r_begin + kRC - 1) / kRC);
    const int nfull = (int)((r_end - r_begin) / kRC);             // chunks with all 16 rows inside the slab
    const int nsub = (nchunks + kSubChunks - 1) / kSubChunks;

    if (warp >= kTQWarp0) {
        // ------------------------------------------------------------------ Q producers (transpose through registers)
        const int pw = warp - kTQWarp0;
        const int r4 = lane & 3, cq = lane >> 2;
        const int rb4 = pw & 3, cbase = pw >> 2;                  // rows rb4*4.. of the chunk; column blocks cbase + 2j
        const int ncb = (nq + 31) >> 5;
        const long rowoff = r_begin + rb4 * 4 + r4;
        const float *pq = Q + rowoff * ldq + q0 + cbase * 32 + cq * 4;
        bool act[kTQItems], vq[kTQItems], st_ok[kTQItems];
        int off[kTQItems];
#pragma unroll
        for (int j = 0; j < kTQItems; ++j) {
            const int cb = cbase + 2 * j;
            act[j] = cb < ncb;                                    // warp-uniform
            vq[j] = vec4 && act[j] && (q0 + cb * 32 + cq * 4 + 3 < qcols);
            st_ok[j] = act[j] && (cb * 32 + lane < nq);           // after the transpose this lane owns tile row cb*32 + lane
            off[j] = tile_off(cb * 32 + lane, rb4);
        }
        float4 rq[kTQPrefetch][kTQItems];
        int fc = 0;
        auto fetch = [&](float4 (&d)[kTQItems]) {
            const bool fullc = fc < nfull;
            const long row = rowoff + (long)fc * kRC;
#pragma unroll
            for (int j = 0; j < kTQItems; ++j) {
                d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (act[j])
                    d[j] = (fullc && vq[j]) ? __ldg(reinterpret_cast<const float4 *>(pq + 64 * j))
                                            : load4<false>(Q, row, r_end, ldq, q0 + (cbase + 2 * j) * 32 + cq * 4, qcols);
            }
            pq += kRC * ldq;
            ++fc;
        };
        const bool b0 = lane & 1, b1 = lane & 2;
        auto transpose_split_store = [&](float4 x, float *hi, float *lo, bool ok) {
            float v0 = x.x, v1 = x.y, v2 = x.z, v3 = x.w;        // 4 consecutive columns at ONE row (r4)
            {
                const float s0 = b0 ? v0 : v1, s1 = b0 ? v2 : v3;
                const float g0 = __shfl_xor_sync(FULL_MASK, s0, 1), g1 = __shfl_xor_sync(FULL_MASK, s1, 1);
                if (b0) { v0 = g0; v2 = g1; } else { v1 = g0; v3 = g1; }
            }
            {
                const float s0 = b1 ? v0 : v2, s1 = b1 ? v1 : v3;
                const float g0 = __shfl_xor_sync(FULL_MASK, s0, 2), g1 = __shfl_xor_sync(FULL_MASK, s1, 2);
                if (b1) { v0 = g0; v1 = g1; } else { v2 = g0; v3 = g1; }
            }
            if (ok) {                                             // ONE column at 4 consecutive rows: a 16-byte core-matrix row
                float4 hh, ll;
                hh.x = tf32_hi(v0); hh.y = tf32_hi(v1); hh.z = tf32_hi(v2); hh.w = tf32_hi(v3);
                ll.x = v0 - hh.x; ll.y = v1 - hh.y; ll.z = v2 - hh.z; ll.w = v3 - hh.w;
                *reinterpret_cast<float4 *>(hi) = hh;
                *reinterpret_cast<float4 *>(lo) = ll;
            }
        };
#pragma unroll
        for (int u = 0; u < kTQPrefetch; ++u)
            if (fc < nchunks) fetch(rq[u]);
        int s = 0;
        uint32_t par = 1;
        float *stage = smem;
        for (int left = nchunks; left > 0; left -= kTQPrefetch) {
#pragma unroll
            for (int u = 0; u < kTQPrefetch; ++u) {
                if (u < left) {
                    mbar_wait(&empty[s], par);
                    float *Qhi = stage, *Qlo = stage + kRC * nq;
#pragma unroll
                    for (int j = 0; j < kTQItems; ++j)
                        if (act[j]) transpose_split_store(rq[u][j], Qhi + off[j], Qlo + off[j], st_ok[j]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    if (fc < nchunks) fetch(rq[u]);
                    stage += stage_floats;
                    if (++s == nstages) { s = 0; par ^= 1u; stage = smem; }
                }
            }
        }
    } else if (warp == kTMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer (A operand from TMEM)
        const uint32_t idesc = make_idesc(kM, nq);
        int s = 0;
        uint32_t par = 0;
        for (int i = 0; i < nchunks; ++i) {
            const int sub = i / kSubChunks, ci = i - sub * kSubChunks, b = sub & 1;
            if (ci == 0) {
                mbar_wait(&tempty[b], (uint32_t)(((sub >> 1) & 1) ^ 1));
                tc_fence_after();
            }
            mbar_wait(&full[s], par);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t q_hi = smem_u32(smem + (size_t)s * stage_floats), q_lo = q_hi + kRC * nq * 4;
                const uint32_t a_hi = tmem + kTRingCol + s * 32, a_lo = a_hi + 16;
                const uint32_t d = tmem + b * kTMaxNQ;
#pragma unroll
                for (int ks = 0; ks < kRC / 8; ++ks) {
                    const uint64_t bh = make_desc(q_hi + ks * 256, 128, kSBO);
                    const uint64_t bl = make_desc(q_lo + ks * 256, 128, kSBO);
                    mma_tf32_ts(d, a_hi + ks * 8, bh, idesc, (ci > 0 || ks > 0) ? 1u : 0u);
                    mma_tf32_ts(d, a_hi + ks * 8, bl, idesc, 1);
                    mma_tf32_ts(d, a_lo + ks * 8, bh, idesc, 1);
                }
                mma_commit(&empty[s]);
                if (ci == kSubChunks - 1 || i == nchunks - 1) mma_commit(&tfull[b]);
            }
            __syncwarp();
            if (++s == nstages) { s = 0; par ^= 1u; }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ P producers (registers -> TMEM, no transpose)
        const int quad = (warp - 4) & 3, half = (warp - 4) >> 2;  // TMEM quadrant; rows 8*half.. of every chunk
        const int pcol = p0 + quad * 32 + lane;
        const bool valid = pcol < pcols;
        const bool any_valid = p0 + quad * 32 < pcols;            // warp-uniform: quadrants past the matrix only keep the barrier protocol
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        const float *pp = P + (r_begin + half * 8) * ldp + pcol;
        float rp[kTPPrefetch][8];
        int fc = 0;
        auto fetch = [&](float (&d)[8]) {
            if (any_valid) {
                const long row0 = r_begin + (long)fc * kRC + half * 8;
                if (fc < nfull && valid) {
                    const float *a = pp;
#pragma unroll
                    for (int r = 0; r < 8; ++r) { d[r] = __ldg(a); a += ldp; }
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) d[r] = (valid && row0 + r < r_end) ? __ldg(pp + r * ldp) : 0.f;
                }
            }
            pp += kRC * ldp;
            ++fc;
        };
#pragma unroll
        for (int u = 0; u < kTPPrefetch; ++u)
            if (fc < nchunks) fetch(rp[u]);
        int s = 0;
        uint32_t par = 1;
        for (int left = nchunks; left > 0; left -= kTPPrefetch) {
#pragma unroll
            for (int u = 0; u < kTPPrefetch; ++u) {
                if (u < left) {
                    mbar_wait(&empty[s], par);
                    tc_fence_after();
                    if (any_valid) {
                        float hi[8], lo[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) { hi[r] = tf32_hi(rp[u][r]); lo[r] = rp[u][r] - hi[r]; }
                        const uint32_t ta = tmem + lane_sel + kTRingCol + s * 32 + half * 8;
                        tmem_st8(ta, hi);
                        tmem_st8(ta + 16, lo);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                    if (fc < nchunks) fetch(rp[u]);
                    if (++s == nstages) { s = 0; par ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: add finished sub-slabs into dW
        const int pcol = p0 + warp * 32 + lane;
        const bool valid = pcol < pcols;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        for (int sub = 0; sub < nsub; ++sub) {
            const int b = sub & 1;
            mbar_wait_sleep(&tfull[b], (uint32_t)((sub >> 1) & 1));
            tc_fence_after();
            for (int c0 = 0; c0 < nq; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + lane_sel + b * kTMaxNQ + c0, v);
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int qc = q0 + c0 + i;
                        if (c0 + i < nq && qc < qcols) {
                            float *dst = transposed ? dW + (long)qc * ldw + pcol : dW + (long)pcol * ldw + qc;
                            atomicAdd(dst, v[i]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// dW[m][n] = X^T.Y with P = the operand that goes through TMEM (<= 128 columns per tile), Q = the transposed one
inline int launch_wgrad_tm(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW,
                           long ldw, bool zeroed, cudaStream_t st) {
    // P: the wider operand if both fit in 128 columns (keeps the transposed operand and its smem traffic small);
    // otherwise whichever fits; if neither fits, Y in tiles of 128
    const bool p_is_y = (m <= kM && n <= kM) ? (n >= m) : (n <= kM ? true : (m <= kM ? false : true));
    const float *Pp = p_is_y ? Y : X, *Qp = p_is_y ? X : Y;
    const long ldp = p_is_y ? ldy : ldx, ldq = p_is_y ? ldx : ldy;
    const int pcols = p_is_y ? n : m, qcols = p_is_y ? m : n;
    const int q16 = (qcols + 15) / 16 * 16;
    const int nq = q16 <= kTMaxNQ ? q16 : 128;
    const int ptiles = (pcols + kM - 1) / kM, qtiles = (qcols + nq - 1) / nq;
    const bool vec = (ldq % 4 == 0) && ((reinterpret_cast<uintptr_t>(Qp) & 15u) == 0);
    const size_t stage_bytes = sizeof(float) * 2 * kRC * (size_t)nq;
    const int nstages = kTMaxStages;
    const size_t smem = stage_bytes * nstages;
    long want = sm_count() / ((long)ptiles * qtiles);
    if (want < 1) want = 1;
    long slab = (rows + want - 1) / want;
    if (slab < 256) slab = 256;
    slab = (slab + kRC - 1) / kRC * kRC;
    const long slabs = (rows + slab - 1) / slab;
    dim3 grid((unsigned)slabs, (unsigned)ptiles, (unsigned)qtiles);
    cudaError_t e = zeroed ? cudaSuccess : cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)m * ldw, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_wgrad_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gemm_wgrad_tm_kernel<<<grid, kTThreads, smem, st>>>(rows, slab, nstages, Pp, ldp, pcols, Qp, ldq, qcols, nq, dW, ldw,
                                                        p_is_y ? 1 : 0, vec ? 1 : 0);
    return pasnl_launch_status();
}

template <int NT>
int launch_wgrad(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW, long ldw,
                 bool zeroed, cudaStream_t st) {
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
    cudaError_t e = zeroed ? cudaSuccess : cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)m * ldw, st);
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

// ---------------------------------------------------------------- weight gradient with one NARROW operand (<= 16 columns)
// dW = X^T.Y where X (xyz / normals: 3, 6, 9 channels) or Y has a handful of columns.  A 128-wide UMMA tile would be
// >90 % padding and cuBLAS runs these at ~1.2 TB/s; the job is one streaming pass over the wide operand, so it is plain
// fp32 SIMT: thread = (4 columns of the wide operand, one of 16 row lanes), the narrow operand's row is a broadcast load.
//   out[s * os + w * ow] += sum_r S[r][s] * W[r][w]
template <int MP, int WQ>
__global__ void __launch_bounds__(256) wgrad_narrow_kernel(long rows, long slab, const float *__restrict__ S, long lds,
                                                           int scols, const float *__restrict__ W, long ldw_, int wcols,
                                                           float *__restrict__ out, long os, long ow, int vec4) {
    // WQ = float4 column groups per row of the wide operand (4, 8 or 16): 256 / WQ rows are in flight per CTA
    constexpr int RL = 256 / WQ;
    __shared__ float red[8][MP][WQ * 4];
    const int t = threadIdx.x, wq = t % WQ, rl = t / WQ, warp = t >> 5;
    const int w0 = blockIdx.y * (WQ * 4) + wq * 4;
    const long r0 = (long)blockIdx.x * slab, r1 = min(rows, r0 + slab);
    float acc[MP][4];
#pragma unroll
    for (int i = 0; i < MP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const bool wvec = vec4 && (w0 + 3 < wcols);
    constexpr int U = 2;
    for (long rb = r0 + rl; rb < r1; rb += RL * U) {
        float4 wv[U];
        float sv[U][MP];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + RL * u;
            wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r1) {
                if (wvec) wv[u] = __ldg(reinterpret_cast<const float4 *>(W + r * ldw_ + w0));
                else wv[u] = load4<false>(W, r, r1, ldw_, w0, wcols);
            }
#pragma unroll
            for (int i = 0; i < MP; ++i) sv[u][i] = (i < scols && r < r1) ? __ldg(S + r * lds + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                acc[i][0] = fmaf(sv[u][i], wv[u].x, acc[i][0]);
                acc[i][1] = fmaf(sv[u][i], wv[u].y, acc[i][1]);
                acc[i][2] = fmaf(sv[u][i], wv[u].z, acc[i][2]);
                acc[i][3] = fmaf(sv[u][i], wv[u].w, acc[i][3]);
            }
    }
    // the row lanes of a warp, then the 8 warps through shared memory, then one atomic per output and CTA
#pragma unroll
    for (int off = WQ; off < 32; off <<= 1)
#pragma unroll
        for (int i = 0; i < MP; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][e] += __shfl_xor_sync(FULL_MASK, acc[i][e], off);
    if ((t & 31) < WQ) {
#pragma unroll
        for (int i = 0; i < MP; ++i)
            *reinterpret_cast<float4 *>(&red[warp][i][wq * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    for (int o = t; o < scols * (WQ * 4); o += 256) {
        const int i = o / (WQ * 4), wc = o % (WQ * 4);
        const int w = blockIdx.y * (WQ * 4) + wc;
        if (w < wcols) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += red[k][i][wc];
            atomicAdd(out + i * os + w * ow, v);
        }
    }
}

inline int launch_wgrad_narrow(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy, float *dW,
                               long ldw, bool zeroed, cudaStream_t st) {
    const bool s_is_x = m <= n;                                   // the narrow operand
    const float *S = s_is_x ? X : Y, *Wd = s_is_x ? Y : X;
    const long lds = s_is_x ? ldx : ldy, ldwide = s_is_x ? ldy : ldx;
    const int scols = s_is_x ? m : n, wcols = s_is_x ? n : m;
    const long os = s_is_x ? ldw : 1, ow = s_is_x ? 1 : ldw;
    const bool vec = (ldwide % 4 == 0) && ((reinterpret_cast<uintptr_t>(Wd) & 15u) == 0);
    const int wq = wcols <= 16 ? 4 : (wcols <= 32 ? 8 : 16);      // float4 column groups per row and CTA
    const int wtiles = (wcols + wq * 4 - 1) / (wq * 4);
    long want = (4L * sm_count() + wtiles - 1) / wtiles;
    long slab = (rows + want - 1) / want;
    if (slab < 256) slab = 256;
    slab = (slab + 31) / 32 * 32;
    dim3 grid((unsigned)((rows + slab - 1) / slab), (unsigned)wtiles);
    cudaError_t e = zeroed ? cudaSuccess : cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)m * ldw, st);
    if (e != cudaSuccess) return (int)e;
#define PASNL_NARROW_WQ(MP, WQ) \
    wgrad_narrow_kernel<MP, WQ><<<grid, 256, 0, st>>>(rows, slab, S, lds, scols, Wd, ldwide, wcols, dW, os, ow, vec ? 1 : 0)
#define PASNL_NARROW(MP)                      \
    do {                                      \
        if (wq == 4) PASNL_NARROW_WQ(MP, 4);  \
        else if (wq == 8) PASNL_NARROW_WQ(MP, 8); \
        else PASNL_NARROW_WQ(MP, 16);         \
    } while (0)
    if (scols <= 4) PASNL_NARROW(4);
    else if (scols <= 8) PASNL_NARROW(8);
    else if (scols <= 12) PASNL_NARROW(12);
    else PASNL_NARROW(16);
#undef PASNL_NARROW
#undef PASNL_NARROW_WQ
    return pasnl_launch_status();
}

}  // namespace gtc

extern "C" int pasnl_gemm_rows_tf32x3(long rows, int k, int n, const float *A, long lda, const float *Bt, long ldb,
                                      float *C, long ldc, double *col_stats, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && k >= 1 && n >= 1 && lda >= k && ldb >= k && ldc >= n);
    if (rows == 0) return PASNL_OK;
    PASNL_CHECK_ARG(A && Bt && C && rows / 128 < 2147483647L && n <= 256 * 65535);
    cudaStream_t st = (cudaStream_t)stream;
    // tall layers whose weights fit one TMEM-resident tile: TMA-fed, weights-stationary kernel (gemm_ws.cu)
    if (pasnl_gemm_rows_ws_eligible(rows, k, n, A, lda) || pasnl_gemm_rows_ws_wide_eligible(rows, k, n, A, lda))
        return pasnl_gemm_rows_ws(rows, k, n, A, lda, Bt, ldb, 0, C, ldc, col_stats, nullptr, nullptr, 0, stream);
    // mid-size layers with TMA-legal operands: both operands by TMA, any k (gemm_tma.cu)
    if (pasnl_gemm_rows_tma_eligible(rows, k, n, A, lda, Bt, ldb))
        return pasnl_gemm_rows_tma(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, stream);
    // many row tiles per SM and small weights: persistent kernel with the weights resident in shared memory
    // (one launch covers up to 160 output columns, so the 134-wide grouped-feature layers stay single-pass)
    const int kpad = (k + gtc::kKC - 1) / gtc::kKC * gtc::kKC;
    if (gtc::persist_enabled() && rows >= 2L * gtc::kM * gtc::sm_count() && n <= 160) {
        const int nt = n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 160));
        if (gtc::persist_stages(kpad, nt) >= 2)
            return gtc::launch_persist_cols(rows, k, n, 0, n, A, lda, Bt, ldb, C, ldc, col_stats, st);
    }
    // UMMA N must be a multiple of 16 (M = 128): the tile is padded in shared memory, stores are bounds-checked.
    // Tile width: as wide as n allows (each CTA splits its A rows once per output tile), but a layer with few row tiles
    // (rows = 4096: the classifier heads, the deep seg layers) takes narrower tiles until the grid covers the SMs.
    // contractions longer than 1024 run as side-by-side <= 1024-wide slabs that add into a zeroed C (split-k)
    const int kslab = k > 1024 ? 1024 : k;
    const int nz = (k + kslab - 1) / kslab;
    if (nz > 1) {
        PASNL_CHECK_ARG(col_stats == nullptr);
        cudaError_t me = cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)n, (size_t)rows, st);
        if (me != cudaSuccess) return (int)me;
    }
    int nt = n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256));
    const long row_tiles = (rows + gtc::kM - 1) / gtc::kM;
    while (nt > gtc::min_tile_width() && row_tiles * ((n + nt - 1) / nt) * nz < gtc::sm_count() * 3 / 4) nt >>= 1;
    if (const char *e = getenv("PASNL_ROWS_NT")) {            // measurement override (tools/sweep_small_gemm.py)
        const int v = atoi(e);
        if (v == 32 || v == 64 || v == 128 || v == 256) nt = v;
    }
    if (nt == 32) return gtc::launch<32>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, kslab, st);
    if (nt == 64) return gtc::launch<64>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, kslab, st);
    if (nt == 128) return gtc::launch<128>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, kslab, st);
    return gtc::launch<256>(rows, k, n, A, lda, Bt, ldb, C, ldc, col_stats, kslab, st);
}

extern "C" int pasnl_gemm_wgrad_tf32x3(long rows, int m, int n, const float *X, long ldx, const float *Y, long ldy,
                                       float *dW, long ldw, int dw_zeroed, void *stream) {
    PASNL_CHECK_ARG(rows >= 0 && m >= 1 && n >= 1 && ldx >= m && ldy >= n && ldw >= n);
    PASNL_CHECK_ARG(dW && (rows == 0 || (X && Y)) && ldw == n && m <= 128 * 65535 && n <= 256 * 65535);
    cudaStream_t st = (cudaStream_t)stream;
    const bool zeroed = dw_zeroed != 0;
    if (rows == 0) {
        cudaError_t e = zeroed ? cudaSuccess : cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)m * ldw, st);
        return (int)e;
    }
    if ((m <= 16 || n <= 16) && rows >= 1024) return gtc::launch_wgrad_narrow(rows, m, n, X, ldx, Y, ldy, dW, ldw, zeroed, st);
    // TMA-legal operands: both streamed as MN-major tiles, no transposes (gemm_wgrad_tma.cu)
    if (pasnl_gemm_wgrad_tma_eligible(rows, m, n, X, ldx, Y, ldy))
        return pasnl_gemm_wgrad_tma(rows, m, n, X, ldx, Y, ldy, dW, ldw, dw_zeroed, nullptr, nullptr, 0, stream);
    // long reductions: one warp-specialised CTA per SM with one operand fed through TMEM; short ones keep the
    // several-CTAs-per-SM kernel
    const int ntile = n <= 32 ? 32 : (n <= 64 ? 64 : 128);
    const long tiles = (long)((m + gtc::kM - 1) / gtc::kM) * ((n + ntile - 1) / ntile);
    if (gtc::persist_enabled() && gtc::wgrad_tm_enabled() && rows * tiles >= 256L * gtc::sm_count()) {
        return gtc::launch_wgrad_tm(rows, m, n, X, ldx, Y, ldy, dW, ldw, zeroed, st);
    }
    if (n <= 32) return gtc::launch_wgrad<32>(rows, m, n, X, ldx, Y, ldy, dW, ldw, zeroed, st);
    if (n <= 64) return gtc::launch_wgrad<64>(rows, m, n, X, ldx, Y, ldy, dW, ldw, zeroed, st);
    return gtc::launch_wgrad<128>(rows, m, n, X, ldx, Y, ldy, dW, ldw, zeroed, st);
}
