// nonlocal_tc.cu -- Non-Local attention core forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as nl_fwd_kernel in nonlocal.cu (reference utils/pointasnl_util.py:195-212): out = softmax(Q Kf^T * scale) V
// without materialising the (S x N) map, plus the row log-sum-exp for the backward.  The parity bar is 1e-4 relative
// against the fp32 reference, which a single TF32 pass (10-bit mantissa) misses by ~10x, so every product is a
// 3xTF32 split:  a = a_hi + a_lo (a_hi = a truncated to TF32, a_lo = a - a_hi),  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi
// (the dropped a_lo.b_lo term is 2^-22 relative).  Three tcgen05.mma.kind::tf32 per k-step, fp32 accumulation in TMEM.
//
// One CTA = 128 threads = 128 query rows (UMMA M = 128: TMEM lane i holds row i, so thread i owns query row i and the
// softmax is thread-local: no shuffles).  Per 64-key tile:
//   1. all threads load the K / V rows from global, split them, and write the four B operands to shared memory in the
//      canonical no-swizzle K-major UMMA layout  [row/8][k/4][row%8][4 floats]  (core matrix = 8 rows x 16 bytes,
//      LBO = 128, SBO = (k/4)*128), filled by warps with coalesced global reads;
//   2. one thread issues S = Q.K^T (3 MMAs per 8-wide k-step) into TMEM columns [0,64), tcgen05.commit -> mbarrier;
//   3. every thread tcgen05.ld's its row of S, does the online-softmax update, writes P (hi/lo) as the next A operand;
//   4. one thread issues PV = P.V into TMEM columns [64, 64+D); every thread adds its PV row into the fp32 output row it
//      keeps in registers (rescaled by the running-max correction).
// This revision is synchronous inside a CTA (no warp-specialised producer yet); smem: D=32 128 KB, D=64 192 KB.
#include "common.cuh"

using namespace pasnl;

namespace tc {

constexpr int kM = 128;     // query rows per CTA (UMMA M)
constexpr int kBN = 64;     // keys per tile (UMMA N of the S product, K extent of the PV product)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);            // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;      // leading (K-direction) byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;      // stride (M/N-direction) byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version 1 (Blackwell)
    return d;                                               // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M x N (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 columns of this thread's TMEM lane
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

// Operand tiles use the canonical no-swizzle K-major layout, row-group major: [row/8][kc = k/4][row%8][4 floats]
// (core matrix = 8 rows x 16 B; LBO = 128 B to the next k core matrix, SBO = KCH*128 B to the next 8-row group, KCH =
// k-chunks per row).  A warp can then fill 8 rows x 4 k-chunks from 8 x 64 contiguous global bytes without conflicts.
__device__ __forceinline__ int tile_off(int kch, int row, int kc) { return ((row >> 3) * kch + kc) * 32 + (row & 7) * 4; }

// write x (split into hi / lo) at operand position (row, kc) of a tile pair with KCH k-chunks per row
__device__ __forceinline__ void store_split4(float *hi, float *lo, int kch, int row, int kc, float4 x) {
    // hi = x with the 13 low mantissa bits cleared (exactly a TF32 value), lo = x - hi (exact in fp32, |lo| < 2^-10 |x|).
    // One LOP + one FADD per element: the F2F conversions (cvt.rna.tf32) run on the 16-lane XU pipe and were the
    // bottleneck of the first revision.  The tensor core drops lo's bits below TF32: 2^-20 relative to x.
    float4 h, l;
    h.x = tf32_hi(x.x); h.y = tf32_hi(x.y); h.z = tf32_hi(x.z); h.w = tf32_hi(x.w);
    l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
    const int o = tile_off(kch, row, kc);
    *reinterpret_cast<float4 *>(hi + o) = h;
    *reinterpret_cast<float4 *>(lo + o) = l;
}

template <int D>
__global__ void __launch_bounds__(kM) nl_fwd_tc_kernel(int s, int n, const float *__restrict__ q,
                                                       const float *__restrict__ kv, float scale_log2,
                                                       float *__restrict__ out, float *__restrict__ lse) {
    extern __shared__ __align__(128) float smem[];
    float *Qhi = smem, *Qlo = Qhi + kM * D;                 // A of S:  128 rows x D
    float *Khi = Qlo + kM * D, *Klo = Khi + kBN * D;        // B of S:  64 rows (keys) x D
    float *Vhi = Klo + kBN * D, *Vlo = Vhi + kBN * D;       // B of PV: D rows (channels) x 64 (keys)
    float *Phi = Vlo + kBN * D, *Plo = Phi + kM * kBN;      // A of PV: 128 rows x 64 (keys)
    __shared__ uint64_t bar_s, bar_pv;
    __shared__ uint32_t tmem_base_slot;

    const int t = threadIdx.x, warp = t >> 5;
    const int bi = blockIdx.y, row0 = blockIdx.x * kM;
    const int row = row0 + t;                               // this thread's query row
    const float *qb = q + (size_t)bi * s * D;
    const float *kvb = kv + (size_t)bi * n * 2 * D;

    if (t == 0) {
        mbar_init(&bar_s, 1);
        mbar_init(&bar_pv, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Q tile (128 rows x D): warp-coalesced, lane = (row%8) + 8*(kc%4); zero rows past s
    {
        const int lane = t & 31, r8 = lane & 7, kq = lane >> 3;
        constexpr int ITEMS = (kM / 8) * (D / 16);
        for (int item = warp; item < ITEMS; item += 4) {
            const int rg = item % (kM / 8), h = item / (kM / 8);
            const int r = rg * 8 + r8, kc = kq + 4 * h;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < s) x = __ldg(reinterpret_cast<const float4 *>(qb + (size_t)(row0 + r) * D + kc * 4));
            store_split4(Qhi, Qlo, D / 4, r, kc, x);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t tmem_s = tmem, tmem_pv = tmem + kBN;                 // column offsets
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;              // this warp's 32 TMEM lanes

    constexpr uint32_t idesc_s = make_idesc(kM, kBN);
    constexpr uint32_t idesc_pv = make_idesc(kM, D);

    float o_acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint32_t phase = 0;

    for (int col0 = 0; col0 < n; col0 += kBN) {
        const int cv = min(kBN, n - col0);
        // ---- 1. K / V tile -> split -> operand layouts (64 keys x D each), warp-coalesced like the Q tile
        {
            const int lane = t & 31, r8 = lane & 7, kq = lane >> 3;
            constexpr int ITEMS = (kBN / 8) * (D / 16);
#pragma unroll
            for (int j = 0; j < ITEMS / 4; ++j) {
                const int item = warp + 4 * j;
                const int rg = item % (kBN / 8), h = item / (kBN / 8);
                const int key = rg * 8 + r8, kc = kq + 4 * h;
                float4 xk = make_float4(0.f, 0.f, 0.f, 0.f), xv = xk;
                if (key < cv) {
                    const float *kr = kvb + (size_t)(col0 + key) * 2 * D + kc * 4;
                    xk = __ldg(reinterpret_cast<const float4 *>(kr));
                    xv = __ldg(reinterpret_cast<const float4 *>(kr + D));
                }
                store_split4(Khi, Klo, D / 4, key, kc, xk);
                // V is the B operand of PV: rows = d (N), k = key -> element (d, key) at tile_off(16, d, key/4) + key%4
                const float vv[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float hh = tf32_hi(vv[i]);
                    const int o = tile_off(kBN / 4, kc * 4 + i, key >> 2) + (key & 3);
                    Vhi[o] = hh;
                    Vlo[o] = vv[i] - hh;
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 2. S = Q K^T
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < D / 8; ++ks) {
                const uint64_t ah = make_desc(smem_u32(Qhi) + ks * 256, 128, (D / 4) * 128);
                const uint64_t al = make_desc(smem_u32(Qlo) + ks * 256, 128, (D / 4) * 128);
                const uint64_t bh = make_desc(smem_u32(Khi) + ks * 256, 128, (D / 4) * 128);
                const uint64_t bl = make_desc(smem_u32(Klo) + ks * 256, 128, (D / 4) * 128);
                mma_tf32(tmem_s, ah, bh, idesc_s, ks > 0);
                mma_tf32(tmem_s, ah, bl, idesc_s, 1);
                mma_tf32(tmem_s, al, bh, idesc_s, 1);
            }
            mma_commit(&bar_s);
        }
        mbar_wait(&bar_s, phase);
        tc_fence_after();
        // ---- 3. online softmax on this thread's row
        float sv[kBN];
        {
            float a[32], b2[32];
            tmem_ld32(tmem_s + lane_sel, a);
            tmem_ld32(tmem_s + lane_sel + 32, b2);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                sv[i] = a[i];
                sv[32 + i] = b2[i];
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kBN; ++j) {
            sv[j] = (j < cv) ? sv[j] * scale_log2 : -INFINITY;
            mx = fmaxf(mx, sv[j]);
        }
        const float m_new = fmaxf(m_run, mx);
        const float alpha = exp2f(m_run - m_new);
        float rs = 0.f;
#pragma unroll
        for (int j = 0; j < kBN; ++j) {
            sv[j] = exp2f(sv[j] - m_new);
            rs += sv[j];
        }
        l_run = fmaf(l_run, alpha, rs);
        m_run = m_new;
#pragma unroll
        for (int kc = 0; kc < kBN / 4; ++kc)
            store_split4(Phi, Plo, kBN / 4, t, kc, make_float4(sv[kc * 4], sv[kc * 4 + 1], sv[kc * 4 + 2], sv[kc * 4 + 3]));
#pragma unroll
        for (int c = 0; c < D; ++c) o_acc[c] *= alpha;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 4. PV = P V
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kBN / 8; ++ks) {
                const uint64_t ah = make_desc(smem_u32(Phi) + ks * 256, 128, (kBN / 4) * 128);
                const uint64_t al = make_desc(smem_u32(Plo) + ks * 256, 128, (kBN / 4) * 128);
                const uint64_t bh = make_desc(smem_u32(Vhi) + ks * 256, 128, (kBN / 4) * 128);
                const uint64_t bl = make_desc(smem_u32(Vlo) + ks * 256, 128, (kBN / 4) * 128);
                mma_tf32(tmem_pv, ah, bh, idesc_pv, ks > 0);
                mma_tf32(tmem_pv, ah, bl, idesc_pv, 1);
                mma_tf32(tmem_pv, al, bh, idesc_pv, 1);
            }
            mma_commit(&bar_pv);
        }
        mbar_wait(&bar_pv, phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
            float pv[32];
            tmem_ld32(tmem_pv + lane_sel + c0, pv);
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[c0 + i] += pv[i];
        }
        phase ^= 1u;
        // the next iteration overwrites K/V/P shared memory and the S/PV TMEM tiles: all MMAs reading them have
        // completed (both commits were waited on); order the TMEM loads above before the next MMA issue
        tc_fence_before();
        __syncthreads();
    }
    if (row < s) {
        const float inv = 1.f / l_run;
        float *o = out + ((size_t)bi * s + row) * D;
#pragma unroll
        for (int c = 0; c < D; c += 4)
            *reinterpret_cast<float4 *>(o + c) =
                make_float4(o_acc[c] * inv, o_acc[c + 1] * inv, o_acc[c + 2] * inv, o_acc[c + 3] * inv);
        lse[(size_t)bi * s + row] = (m_run + log2f(l_run)) * kLn2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
    }
}

template <int D>
int launch(int b, int s, int n, const float *q, const float *kv, float scale, float *out, float *lse, cudaStream_t st) {
    auto kern = nl_fwd_tc_kernel<D>;
    const size_t smem = sizeof(float) * (size_t)(2 * kM * D + 4 * kBN * D + 2 * kM * kBN);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div_i(s, kM), (unsigned)b);
    kern<<<grid, kM, smem, st>>>(s, n, q, kv, scale * kLog2e, out, lse);
    return pasnl_launch_status();
}

}  // namespace tc

// Called by pasnl_nl_core_fwd (nonlocal.cu) for the shapes where a 128 x 64 x D tile fills the tensor pipe.
int nl_core_fwd_tcgen05(int b, int s, int n, int d, const float *q, const float *kv, float scale, float *out,
                         float *lse, cudaStream_t st) {
    if (d == 32) return tc::launch<32>(b, s, n, q, kv, scale, out, lse, st);
    if (d == 64) return tc::launch<64>(b, s, n, q, kv, scale, out, lse, st);
    return PASNL_EINVAL;
}
