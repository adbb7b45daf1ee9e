// nonlocal_bwd_tc.cu -- Non-Local attention core BACKWARD on tcgen05 tensor cores (3xTF32, TMEM accumulators), sm_100a.
//
// Contract of nl_bwd_kernel in nonlocal.cu (gradient of out = softmax(Q Kf^T * scale) V, reference
// utils/pointasnl_util.py:195-212 under TF autodiff): with P = exp(S*scale - lse) recomputed from the saved row
// log-sum-exp and delta[q] = sum_c dO[q][c] * O[q][c],
//     dP = dO V^T,   dS = P o (dP - delta) * scale,   dQ = dS Kf,   dKf = dS^T Q,   dV = P^T dO.
// A UMMA operand must have its contraction index contiguous, and an accumulator row lives in one thread (TMEM lane), so
// the work is split by which index is the accumulator row -- two kernels, no atomics between them:
//   * nl_bwd_dq_tc_kernel: CTA = 128 QUERIES.  Per 64-key tile: S = Q.Kf^T and dP = dO.V^T (all operands in their
//     natural row-major form), thread-local P / dS, then dQ += dS.Kf (Kf staged transposed, like V in the forward).
//   * nl_bwd_dkv_tc_kernel: CTA = 128 KEYS.  Per 32-query tile: S' = Kf.Q^T and dP' = V.dO^T give the TRANSPOSED maps
//     with the key as the accumulator row, so P' and dS' are already the A operands of dV += P'.dO and dKf += dS'.Q
//     (Q and dO tiles staged transposed); the per-query lse / delta are broadcast from shared memory.
// Seven 3xTF32 GEMMs per (query tile, key tile) pair instead of the five of a fused kernel with atomics -- tensor time
// is not what bounds this op.  Output accumulators stay in TMEM across tiles and are drained into registers every 1024
// contraction rows (the TMEM adder truncates: ~7e-9 relative per accumulated row).
// D = 32 (cls layer 1, every seg / KITTI layer that is large enough to matter); other widths keep the fp32 FMA kernel.
#include "common.cuh"

using namespace pasnl;

namespace tcb {

constexpr int kM = 128;
constexpr int kThreads = 256;   // two threads per accumulator row: warps w and w+4 share TMEM lanes 32*(w%4).. and split the columns
constexpr float kLog2e = 1.4426950408889634f;

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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// canonical no-swizzle K-major operand tile, row-group major: [row/8][kc = k/4][row%8][4 floats]; KCH = k-chunks per row
__device__ __forceinline__ int tile_off(int kch, int row, int kc) { return ((row >> 3) * kch + kc) * 32 + (row & 7) * 4; }
__device__ __forceinline__ void store_split4(float *hi, float *lo, int kch, int row, int kc, float4 x) {
    float4 h, l;
    h.x = tf32_hi(x.x); h.y = tf32_hi(x.y); h.z = tf32_hi(x.z); h.w = tf32_hi(x.w);
    l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
    const int o = tile_off(kch, row, kc);
    *reinterpret_cast<float4 *>(hi + o) = h;
    *reinterpret_cast<float4 *>(lo + o) = l;
}
// x = 4 consecutive k of row `row`; store it into the TRANSPOSED tile (rows = k index, contraction = row index)
__device__ __forceinline__ void store_split4_t(float *hi, float *lo, int kch_t, int row, int kc, float4 x) {
    const float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float hh = tf32_hi(v[i]);
        const int o = tile_off(kch_t, kc * 4 + i, row >> 2) + (row & 3);
        hi[o] = hh;
        lo[o] = v[i] - hh;
    }
}
// A . B^T for two K-major tile pairs (hi/lo): 3 MMAs per 8-wide k-step
__device__ __forceinline__ void mma3(uint32_t tmem_d, const float *a_hi, const float *a_lo, int a_kch, const float *b_hi,
                                     const float *b_lo, int b_kch, int ksteps, uint32_t idesc, bool accumulate) {
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ah = make_desc(smem_u32(a_hi) + ks * 256, 128, a_kch * 128);
        const uint64_t al = make_desc(smem_u32(a_lo) + ks * 256, 128, a_kch * 128);
        const uint64_t bh = make_desc(smem_u32(b_hi) + ks * 256, 128, b_kch * 128);
        const uint64_t bl = make_desc(smem_u32(b_lo) + ks * 256, 128, b_kch * 128);
        mma_tf32(tmem_d, ah, bh, idesc, (accumulate || ks > 0) ? 1u : 0u);
        mma_tf32(tmem_d, ah, bl, idesc, 1);
        mma_tf32(tmem_d, al, bh, idesc, 1);
    }
}
// rows [row0, row0+ROWS) x D of a row-major matrix (pitch ld) -> split operand tile; rows >= limit are zero
template <int ROWS, int D>
__device__ __forceinline__ void load_tile(float *hi, float *lo, const float *__restrict__ src, long ld, int row0,
                                          int limit, int warp, int lane) {
    const int r8 = lane & 7, kq = lane >> 3;
    constexpr int ITEMS = (ROWS / 8) * (D / 16);
    for (int item = warp; item < ITEMS; item += kThreads / 32) {
        const int rg = item % (ROWS / 8), h = item / (ROWS / 8);
        const int r = rg * 8 + r8, kc = kq + 4 * h;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < limit) x = __ldg(reinterpret_cast<const float4 *>(src + (long)(row0 + r) * ld + kc * 4));
        store_split4(hi, lo, D / 4, r, kc, x);
    }
}
// the same rows, stored as the natural tile AND as the transposed tile (rows = channel, contraction = row index)
template <int ROWS, int D>
__device__ __forceinline__ void load_tile_both(float *hi, float *lo, float *thi, float *tlo, const float *__restrict__ src,
                                               long ld, int row0, int limit, int warp, int lane) {
    const int r8 = lane & 7, kq = lane >> 3;
    constexpr int ITEMS = (ROWS / 8) * (D / 16);
    for (int item = warp; item < ITEMS; item += kThreads / 32) {
        const int rg = item % (ROWS / 8), h = item / (ROWS / 8);
        const int r = rg * 8 + r8, kc = kq + 4 * h;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < limit) x = __ldg(reinterpret_cast<const float4 *>(src + (long)(row0 + r) * ld + kc * 4));
        if (hi != nullptr) store_split4(hi, lo, D / 4, r, kc, x);
        store_split4_t(thi, tlo, ROWS / 4, r, kc, x);
    }
}

// The same tile traffic split in two so that the global loads of the NEXT tile can be issued before the MMAs of the
// current one are waited for: fetch (global -> registers) and store (registers -> split operand tiles).
template <int ROWS, int D>
struct TileRegs {
    static constexpr int ITEMS = (ROWS / 8) * (D / 16);
    static constexpr int PER_WARP = (ITEMS + kThreads / 32 - 1) / (kThreads / 32);
    float4 v[PER_WARP];
};
template <int ROWS, int D>
__device__ __forceinline__ void fetch_tile(TileRegs<ROWS, D> &tr, const float *__restrict__ src, long ld, int row0, int limit,
                                           int warp, int lane) {
    const int r8 = lane & 7, kq = lane >> 3;
#pragma unroll
    for (int j = 0; j < TileRegs<ROWS, D>::PER_WARP; ++j) {
        const int item = warp + (kThreads / 32) * j;
        const int rg = item % (ROWS / 8), h = item / (ROWS / 8);
        const int r = rg * 8 + r8, kc = kq + 4 * h;
        tr.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < TileRegs<ROWS, D>::ITEMS && row0 + r < limit)
            tr.v[j] = __ldg(reinterpret_cast<const float4 *>(src + (long)(row0 + r) * ld + kc * 4));
    }
}
// natural tile (hi/lo, may be null) and/or transposed tile (thi/tlo, may be null)
template <int ROWS, int D>
__device__ __forceinline__ void store_tile(const TileRegs<ROWS, D> &tr, float *hi, float *lo, float *thi, float *tlo, int warp,
                                           int lane) {
    const int r8 = lane & 7, kq = lane >> 3;
#pragma unroll
    for (int j = 0; j < TileRegs<ROWS, D>::PER_WARP; ++j) {
        const int item = warp + (kThreads / 32) * j;
        if (item < TileRegs<ROWS, D>::ITEMS) {
            const int rg = item % (ROWS / 8), h = item / (ROWS / 8);
            const int r = rg * 8 + r8, kc = kq + 4 * h;
            if (hi != nullptr) store_split4(hi, lo, D / 4, r, kc, tr.v[j]);
            if (thi != nullptr) store_split4_t(thi, tlo, ROWS / 4, r, kc, tr.v[j]);
        }
    }
}

constexpr int kFlushRows = 1024;   // contraction rows accumulated in TMEM before the partial is drained into registers

// ------------------------------------------------------------------------------------------------ dQ (CTA = 128 queries)
template <int D, int BN>
__global__ void __launch_bounds__(kThreads) nl_bwd_dq_tc_kernel(int s, int n, int keys_per_cta, const float *__restrict__ q,
                                                          const float *__restrict__ kv, const float *__restrict__ dout,
                                                          const float *__restrict__ lse, const float *__restrict__ delta,
                                                          float scale, float *__restrict__ grad_q, int use_atomics) {
    extern __shared__ __align__(128) float smem[];
    float *Qhi = smem, *Qlo = Qhi + kM * D;                  // A of S
    float *Ohi = Qlo + kM * D, *Olo = Ohi + kM * D;          // A of dP (dO rows)
    float *Khi = Olo + kM * D, *Klo = Khi + BN * D;          // B of S
    float *Vhi = Klo + BN * D, *Vlo = Vhi + BN * D;          // B of dP
    float *Kthi = Vlo + BN * D, *Ktlo = Kthi + D * BN;       // B of dQ: rows = channel, contraction = key
    float *Shi = Ktlo + D * BN, *Slo = Shi + kM * BN;        // A of dQ: dS, 128 x BN
    __shared__ uint64_t bar1, bar2;
    __shared__ uint32_t tmem_slot;
    constexpr int kTmemCols = 256;
    static_assert(2 * BN + D <= kTmemCols, "TMEM budget");

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int trow = t & (kM - 1), half = t >> 7;            // accumulator row of this thread, which half of the columns
    const int bi = blockIdx.y, row0 = blockIdx.x * kM, row = row0 + trow;
    const int key_begin = blockIdx.z * keys_per_cta, key_end = min(n, key_begin + keys_per_cta);
    const float *qb = q + (size_t)bi * s * D, *dob = dout + (size_t)bi * s * D;
    const float *kvb = kv + (size_t)bi * n * 2 * D;
    const float scale_log2 = scale * kLog2e;

    if (t == 0) {
        mbar_init(&bar1, 1);
        mbar_init(&bar2, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    load_tile<kM, D>(Qhi, Qlo, qb, D, row0, s, warp, lane);
    load_tile<kM, D>(Ohi, Olo, dob, D, row0, s, warp, lane);
    float lse2 = 0.f, del = 0.f;
    if (row < s) {
        lse2 = __ldg(lse + (size_t)bi * s + row) * kLog2e;
        del = __ldg(delta + (size_t)bi * s + row);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmem_s = tmem, tmem_dp = tmem + BN, tmem_dq = tmem + 2 * BN;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    constexpr uint32_t idesc_s = make_idesc(kM, BN), idesc_o = make_idesc(kM, D);
    static_assert(BN == 64 && D == 32, "column split: 32 S columns and 16 dQ columns per thread");

    float acc[D / 2];
#pragma unroll
    for (int c = 0; c < D / 2; ++c) acc[c] = 0.f;
    uint32_t phase = 0;
    int in_group = 0;                                        // key tiles accumulated in TMEM since the last drain
    constexpr int kGroup = kFlushRows / BN;

    TileRegs<BN, D> rk, rv;                                  // the next key tile, in flight while this one is processed
    fetch_tile<BN, D>(rk, kvb, 2 * D, key_begin, key_end, warp, lane);
    fetch_tile<BN, D>(rv, kvb + D, 2 * D, key_begin, key_end, warp, lane);
    for (int col0 = key_begin; col0 < key_end; col0 += BN) {
        // ---- 1. Kf / V tiles: natural (B of S, dP) and Kf transposed (B of dQ); then request the next tile
        store_tile<BN, D>(rk, Khi, Klo, Kthi, Ktlo, warp, lane);
        store_tile<BN, D>(rv, Vhi, Vlo, nullptr, nullptr, warp, lane);
        if (col0 + BN < key_end) {
            fetch_tile<BN, D>(rk, kvb, 2 * D, col0 + BN, key_end, warp, lane);
            fetch_tile<BN, D>(rv, kvb + D, 2 * D, col0 + BN, key_end, warp, lane);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 2. S = Q Kf^T, dP = dO V^T
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            mma3(tmem_s, Qhi, Qlo, D / 4, Khi, Klo, D / 4, D / 8, idesc_s, false);
            mma3(tmem_dp, Ohi, Olo, D / 4, Vhi, Vlo, D / 4, D / 8, idesc_s, false);
            mma_commit(&bar1);
        }
        mbar_wait(&bar1, phase);
        tc_fence_after();
        // ---- 3. this thread's row of P and dS (keys past the end: P = 0)
        {
            const int h = half;
            float sv[32], dp[32];
            tmem_ld32(tmem_s + lane_sel + h * 32, sv);
            tmem_ld32(tmem_dp + lane_sel + h * 32, dp);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float p = (col0 + h * 32 + j < key_end) ? exp2f(fmaf(sv[j], scale_log2, -lse2)) : 0.f;
                sv[j] = p * (dp[j] - del) * scale;
            }
#pragma unroll
            for (int kc = 0; kc < 8; ++kc)
                store_split4(Shi, Slo, BN / 4, trow, h * 8 + kc, make_float4(sv[kc * 4], sv[kc * 4 + 1], sv[kc * 4 + 2], sv[kc * 4 + 3]));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 4. dQ += dS Kf
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            mma3(tmem_dq, Shi, Slo, BN / 4, Kthi, Ktlo, BN / 4, BN / 8, idesc_o, in_group > 0);
            mma_commit(&bar2);
        }
        mbar_wait(&bar2, phase);
        tc_fence_after();
        phase ^= 1u;
        if (++in_group == kGroup || col0 + BN >= key_end) {
            float v[16];
            tmem_ld16(tmem_dq + lane_sel + half * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += v[i];
            in_group = 0;
        }
        tc_fence_before();
        __syncthreads();
    }
    if (row < s) {
        float *g = grad_q + ((size_t)bi * s + row) * D + half * 16;
        if (use_atomics) {
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(g + c, acc[c]);
        } else {
#pragma unroll
            for (int c = 0; c < 16; c += 4)
                *reinterpret_cast<float4 *>(g + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------------ dKf, dV (CTA = 128 keys)
template <int D, int BQ>
__global__ void __launch_bounds__(kThreads) nl_bwd_dkv_tc_kernel(int s, int n, const float *__restrict__ q,
                                                           const float *__restrict__ kv, const float *__restrict__ dout,
                                                           const float *__restrict__ lse, const float *__restrict__ delta,
                                                           float scale, float *__restrict__ grad_kv) {
    extern __shared__ __align__(128) float smem[];
    float *Khi = smem, *Klo = Khi + kM * D;                  // A of S'
    float *Vhi = Klo + kM * D, *Vlo = Vhi + kM * D;          // A of dP'
    float *Qhi = Vlo + kM * D, *Qlo = Qhi + BQ * D;          // B of S'
    float *Ohi = Qlo + BQ * D, *Olo = Ohi + BQ * D;          // B of dP' (dO rows)
    float *Qthi = Olo + BQ * D, *Qtlo = Qthi + D * BQ;       // B of dKf: rows = channel, contraction = query
    float *Othi = Qtlo + D * BQ, *Otlo = Othi + D * BQ;      // B of dV
    float *Phi = Otlo + D * BQ, *Plo = Phi + kM * BQ;        // A of dV: P', 128 x BQ
    float *Shi = Plo + kM * BQ, *Slo = Shi + kM * BQ;        // A of dKf: dS'
    float *lse_c = Slo + kM * BQ, *del_c = lse_c + BQ;       // per-query lse (log2 units) / delta of the tile
    __shared__ uint64_t bar1, bar2;
    __shared__ uint32_t tmem_slot;
    constexpr int kTmemCols = 128;
    static_assert(2 * BQ + 2 * D <= kTmemCols && BQ == 32, "TMEM budget / one tcgen05.ld per map");

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int trow = t & (kM - 1), half = t >> 7;
    const int bi = blockIdx.y, key0 = blockIdx.x * kM, key = key0 + trow;
    const float *qb = q + (size_t)bi * s * D, *dob = dout + (size_t)bi * s * D;
    const float *kvb = kv + (size_t)bi * n * 2 * D;
    const float scale_log2 = scale * kLog2e;

    if (t == 0) {
        mbar_init(&bar1, 1);
        mbar_init(&bar2, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    load_tile<kM, D>(Khi, Klo, kvb, 2 * D, key0, n, warp, lane);
    load_tile<kM, D>(Vhi, Vlo, kvb + D, 2 * D, key0, n, warp, lane);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmem_s = tmem, tmem_dp = tmem + BQ, tmem_dv = tmem + 2 * BQ, tmem_dk = tmem + 2 * BQ + D;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    constexpr uint32_t idesc_s = make_idesc(kM, BQ), idesc_o = make_idesc(kM, D);
    static_assert(D == 32, "column split: 16 columns of every map per thread");

    float acc_k[16], acc_v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc_k[c] = acc_v[c] = 0.f;
    uint32_t phase = 0;
    int in_group = 0;
    constexpr int kGroup = kFlushRows / BQ;

    TileRegs<BQ, D> rq, ro;                                  // the next query tile, in flight while this one is processed
    fetch_tile<BQ, D>(rq, qb, D, 0, s, warp, lane);
    fetch_tile<BQ, D>(ro, dob, D, 0, s, warp, lane);
    float lse_n = 0.f, del_n = 0.f;
    if (t < BQ && t < s) {
        lse_n = __ldg(lse + (size_t)bi * s + t) * kLog2e;
        del_n = __ldg(delta + (size_t)bi * s + t);
    }
    for (int q0 = 0; q0 < s; q0 += BQ) {
        // ---- 1. Q / dO tiles, natural and transposed; per-query lse and delta; then request the next tile
        store_tile<BQ, D>(rq, Qhi, Qlo, Qthi, Qtlo, warp, lane);
        store_tile<BQ, D>(ro, Ohi, Olo, Othi, Otlo, warp, lane);
        if (t < BQ) {
            lse_c[t] = lse_n;
            del_c[t] = del_n;
        }
        if (q0 + BQ < s) {
            fetch_tile<BQ, D>(rq, qb, D, q0 + BQ, s, warp, lane);
            fetch_tile<BQ, D>(ro, dob, D, q0 + BQ, s, warp, lane);
            if (t < BQ) {
                const bool ok = q0 + BQ + t < s;
                lse_n = ok ? __ldg(lse + (size_t)bi * s + q0 + BQ + t) * kLog2e : 0.f;
                del_n = ok ? __ldg(delta + (size_t)bi * s + q0 + BQ + t) : 0.f;
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 2. S' = Kf Q^T, dP' = V dO^T   (key = accumulator row)
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            mma3(tmem_s, Khi, Klo, D / 4, Qhi, Qlo, D / 4, D / 8, idesc_s, false);
            mma3(tmem_dp, Vhi, Vlo, D / 4, Ohi, Olo, D / 4, D / 8, idesc_s, false);
            mma_commit(&bar1);
        }
        mbar_wait(&bar1, phase);
        tc_fence_after();
        // ---- 3. this key's row of P' and dS' (queries past the end have zero dO rows and lse = delta = 0: they add 0)
        {
            float sv[16], dp[16];
            tmem_ld16(tmem_s + lane_sel + half * 16, sv);
            tmem_ld16(tmem_dp + lane_sel + half * 16, dp);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float p = exp2f(fmaf(sv[j], scale_log2, -lse_c[half * 16 + j]));
                sv[j] = p;
                dp[j] = p * (dp[j] - del_c[half * 16 + j]) * scale;
            }
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                store_split4(Phi, Plo, BQ / 4, trow, half * 4 + kc, make_float4(sv[kc * 4], sv[kc * 4 + 1], sv[kc * 4 + 2], sv[kc * 4 + 3]));
                store_split4(Shi, Slo, BQ / 4, trow, half * 4 + kc, make_float4(dp[kc * 4], dp[kc * 4 + 1], dp[kc * 4 + 2], dp[kc * 4 + 3]));
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 4. dV += P' dO, dKf += dS' Q
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            mma3(tmem_dv, Phi, Plo, BQ / 4, Othi, Otlo, BQ / 4, BQ / 8, idesc_o, in_group > 0);
            mma3(tmem_dk, Shi, Slo, BQ / 4, Qthi, Qtlo, BQ / 4, BQ / 8, idesc_o, in_group > 0);
            mma_commit(&bar2);
        }
        mbar_wait(&bar2, phase);
        tc_fence_after();
        phase ^= 1u;
        if (++in_group == kGroup || q0 + BQ >= s) {
            float v[16];
            tmem_ld16(tmem_dv + lane_sel + half * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc_v[i] += v[i];
            tmem_ld16(tmem_dk + lane_sel + half * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc_k[i] += v[i];
            in_group = 0;
        }
        tc_fence_before();
        __syncthreads();
    }
    if (key < n) {
        float *g = grad_kv + ((size_t)bi * n + key) * 2 * D + half * 16;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            *reinterpret_cast<float4 *>(g + c) = make_float4(acc_k[c], acc_k[c + 1], acc_k[c + 2], acc_k[c + 3]);
            *reinterpret_cast<float4 *>(g + D + c) = make_float4(acc_v[c], acc_v[c + 1], acc_v[c + 2], acc_v[c + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

template <int D>
int launch(int b, int s, int n, const float *q, const float *kv, const float *lse, const float *delta, float scale,
           const float *dout, float *gq, float *gkv, cudaStream_t st) {
    constexpr int BN = 64, BQ = 32;
    cudaError_t e;
    {
        auto kern = nl_bwd_dq_tc_kernel<D, BN>;
        const size_t smem = sizeof(float) * (size_t)(4 * kM * D + 6 * BN * D + 2 * kM * BN);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        // few query tiles (seg layer 1: 10 per cloud): split the key range over blockIdx.z and reduce with atomics
        const int qtiles = ceil_div_i(s, kM);
        int splits = 1;
        const int ctas = qtiles * b;
        if (ctas < 148) {
            splits = (2 * 148 + ctas - 1) / ctas;
            const int max_splits = ceil_div_i(n, 4 * BN);
            if (splits > max_splits) splits = max_splits;
            if (splits < 1) splits = 1;
        }
        int keys_per_cta = ceil_div_i(ceil_div_i(n, splits), BN) * BN;
        splits = ceil_div_i(n, keys_per_cta);
        if (splits > 1) {
            e = cudaMemsetAsync(gq, 0, sizeof(float) * (size_t)b * s * D, st);
            if (e != cudaSuccess) return (int)e;
        }
        dim3 grid((unsigned)qtiles, (unsigned)b, (unsigned)splits);
        kern<<<grid, kThreads, smem, st>>>(s, n, keys_per_cta, q, kv, dout, lse, delta, scale, gq, splits > 1 ? 1 : 0);
    }
    {
        auto kern = nl_bwd_dkv_tc_kernel<D, BQ>;
        const size_t smem = sizeof(float) * (size_t)(4 * kM * D + 8 * BQ * D + 4 * kM * BQ + 2 * BQ);
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((unsigned)ceil_div_i(n, kM), (unsigned)b);
        kern<<<grid, kThreads, smem, st>>>(s, n, q, kv, dout, lse, delta, scale, gkv);
    }
    return pasnl_launch_status();
}

}  // namespace tcb

// Called by pasnl_nl_core_bwd (nonlocal.cu) after the delta pass, for d == 32 and shapes that fill 128-row tiles.
int nl_core_bwd_tcgen05(int b, int s, int n, int d, const float *q, const float *kv, const float *lse, const float *delta,
                         float scale, const float *dout, float *gq, float *gkv, cudaStream_t st) {
    if (d == 32) return tcb::launch<32>(b, s, n, q, kv, lse, delta, scale, dout, gq, gkv, st);
    return PASNL_EINVAL;
}
