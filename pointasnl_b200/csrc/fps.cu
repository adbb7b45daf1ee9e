// fps.cu -- farthest point sampling for sm_100a.
//
// Replaces farthestpointsamplingKernel (reference tf_ops/sampling/tf_sampling_g.cu:105-170), which runs
// <<<32,512>>> with the running min-distance array in GLOBAL memory, re-reads xyz from shared/global every round
// and does a 9-level __syncthreads tree per round.  Design here (SURVEY.md 7.3 "FPS latency"):
//   * one CTA (or one thread-block CLUSTER for big clouds) per cloud, persistent over all m rounds;
//   * the cloud is staged once into shared memory with a TMA bulk copy, then every thread keeps its P points AND
//     their running min distances in registers for the whole kernel -- nothing touches global memory in the loop;
//   * per round: register sweep, warp arg-max with two redux.sync (max of the distance bits, min of the tie key),
//     one shared-memory hop + ONE __syncthreads, second redux pair computed redundantly by every warp;
//   * clusters: CTA winners are pushed to every peer's shared memory (DSMEM) and combined after one cluster barrier.
// FPS is latency/sync bound, not HBM bound (compulsory traffic is 12*b*n + 4*b*m bytes in total).
//
// Bit-exact contract: start index 0; running distance init 1e38f;
// d = fmaf(dz,dz,fmaf(dx,dx,dy*dy)) with d* = p_k - p_last; d2 = min(d, running); arg-max over d2 where ties go to
// the lowest (k mod 512) and then the lowest k (the order the reference's 512-thread strided scan + keep-left tree
// produces).  Distances are non-negative so their IEEE bit patterns order like unsigned integers.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;
using namespace pasnl;

namespace {

// tie key: smaller wins.  9 bits of (k mod 512) above 22 bits of (k div 512): valid for k < 2^31.
__device__ __forceinline__ unsigned tie_key(int k) { return ((unsigned)(k & 511) << 22) | ((unsigned)k >> 9); }
__device__ __forceinline__ int tie_key_decode(unsigned t) { return (int)(((t & 0x3fffffu) << 9) | (t >> 22)); }

struct __align__(16) Slot {  // one candidate: 32 bytes
    unsigned bits;           // IEEE bits of the (non-negative) distance
    unsigned tk;             // tie key, 0xffffffff = "no candidate"
    float x, y, z;
    unsigned pad[3];
};

constexpr int kMaxWarps = 32;
constexpr int kMaxCluster = 8;

// ---- DSMEM exchange primitives (cluster path): remote shared-memory stores that complete a transaction count on
// the RECEIVER's mbarrier, so a round needs no cluster-wide barrier -- each CTA just waits for its csize slots.
__device__ __forceinline__ uint32_t map_to_rank(const void *local_smem, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t dst, unsigned a, unsigned b, unsigned c, unsigned d, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "XWAIT:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra XDONE;\n\t"
        "bra XWAIT;\n\t"
        "XDONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Reduce (bits, tk) over the warp: returns true on the single winning lane.
__device__ __forceinline__ bool warp_argmax(unsigned bits, unsigned tk, unsigned &wbits, unsigned &wtk) {
    wbits = __reduce_max_sync(FULL_MASK, bits);
    wtk = __reduce_min_sync(FULL_MASK, bits == wbits ? tk : 0xffffffffu);
    return bits == wbits && tk == wtk && tk != 0xffffffffu;
}

template <int P, int THREADS, bool CLUSTER>
__global__ void __launch_bounds__(THREADS) fps_kernel(int n, int m, int chunk, const float *__restrict__ xyz,
                                                      int *__restrict__ out) {
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Slot wslot[2][NW];            // per-warp winners, double buffered over rounds
    __shared__ Slot cslot[2][kMaxCluster];   // per-CTA winners of the cluster (written by peers through DSMEM)
    __shared__ uint64_t bar;
    __shared__ uint64_t xbar[2];             // receive barriers of the DSMEM exchange, one per slot buffer

    int cloud, rank = 0, csize = 1;
    if constexpr (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        rank = (int)cluster.block_rank();
        csize = (int)cluster.num_blocks();
        cloud = blockIdx.x / csize;
    } else {
        cloud = blockIdx.x;
    }
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *cloud_xyz = xyz + (size_t)cloud * n * 3;
    int *cloud_out = out + (size_t)cloud * m;

    // this CTA's slice [base, base+cnt) of the cloud
    const int base = rank * chunk;
    const int cnt = max(0, min(n - base, chunk));

    // ---- stage the slice into shared memory (TMA bulk copy) and pull this thread's points into registers
    float *buf = reinterpret_cast<float *>(smem_raw);
    if (t == 0) {
        mbar_init(&bar, 1);
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if constexpr (CLUSTER) cg::this_cluster().sync();   // peers' receive barriers exist before anybody sends
    uint32_t xparity = 0;                               // bit pb = phase parity of xbar[pb]
    const float *pts = stage_floats(buf, cloud_xyz + (size_t)base * 3, cnt * 3, &bar);
    stage_wait(&bar, 0);

    float px[P], py[P], pz[P], pd[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int kl = t + i * THREADS;  // local index inside the slice
        const bool valid = kl < cnt;
        px[i] = valid ? pts[kl * 3 + 0] : 0.f;
        py[i] = valid ? pts[kl * 3 + 1] : 0.f;
        pz[i] = valid ? pts[kl * 3 + 2] : 0.f;
        pd[i] = 1e38f;
    }

    // first pick is index 0 (reference :114-116)
    float x1 = __ldg(cloud_xyz + 0), y1 = __ldg(cloud_xyz + 1), z1 = __ldg(cloud_xyz + 2);
    if (t == 0 && rank == 0) cloud_out[0] = 0;

    for (int j = 1; j < m; ++j) {
        const int pb = j & 1;
        // ---- register sweep: update running minima, thread-local arg-max
        unsigned bbits = 0u, btk = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const float d = dist2_fma(px[i] - x1, py[i] - y1, pz[i] - z1);
            const float d2 = fminf(d, pd[i]);
            pd[i] = d2;
            const unsigned bits = __float_as_uint(d2);
            const int kl = t + i * THREADS;
            const unsigned tk = tie_key(base + kl);  // recomputed: cheaper than P more live registers
            const bool better = kl < cnt && (bits > bbits || (bits == bbits && tk < btk));
            bbits = better ? bits : bbits;
            btk = better ? tk : btk;
        }
        // ---- warp level
        unsigned wbits, wtk;
        const bool win = warp_argmax(bbits, btk, wbits, wtk);
        if (lane == 0) {
            wslot[pb][warp].bits = wbits;
            wslot[pb][warp].tk = wtk;
        }
        if (win) {
            const int kl = tie_key_decode(wtk) - base;
            wslot[pb][warp].x = pts[kl * 3 + 0];
            wslot[pb][warp].y = pts[kl * 3 + 1];
            wslot[pb][warp].z = pts[kl * 3 + 2];
        }
        __syncthreads();
        // ---- CTA level: every warp reduces the NW slots redundantly (no second barrier needed)
        unsigned sbits = 0u, stk = 0xffffffffu;
        if (lane < NW) {
            sbits = wslot[pb][lane].bits;
            stk = wslot[pb][lane].tk;
        }
        unsigned cbits, ctk;
        warp_argmax(sbits, stk, cbits, ctk);
        int src = __ffs(__ballot_sync(FULL_MASK, lane < NW && sbits == cbits && stk == ctk)) - 1;
        src = max(src, 0);

        if constexpr (CLUSTER) {
            // warp 0 pushes this CTA's winner into cslot[pb][rank] of EVERY CTA of the cluster (itself included)
            // with st.async; each receiver's mbarrier completes when all csize x 16 bytes x 2 have landed.
            if (t == 0) mbar_expect_tx(&xbar[pb], (uint32_t)csize * 32u);
            if (warp == 0 && lane < csize) {
                const uint32_t dst = map_to_rank(&cslot[pb][rank], lane);
                const uint32_t rbar = map_to_rank(&xbar[pb], lane);
                st_async_v4(dst, cbits, ctk, __float_as_uint(wslot[pb][src].x), __float_as_uint(wslot[pb][src].y), rbar);
                st_async_v4(dst + 16, __float_as_uint(wslot[pb][src].z), 0u, 0u, 0u, rbar);
            }
            mbar_wait_cluster(&xbar[pb], (xparity >> pb) & 1u);
            xparity ^= 1u << pb;
            unsigned gbits = 0u, gtk = 0xffffffffu;
            if (lane < csize) {
                gbits = cslot[pb][lane].bits;
                gtk = cslot[pb][lane].tk;
            }
            unsigned fbits, ftk;
            warp_argmax(gbits, gtk, fbits, ftk);
            int csrc = __ffs(__ballot_sync(FULL_MASK, lane < csize && gbits == fbits && gtk == ftk)) - 1;
            csrc = max(csrc, 0);
            x1 = cslot[pb][csrc].x;
            y1 = cslot[pb][csrc].y;
            z1 = cslot[pb][csrc].z;
            if (t == 0 && rank == 0) cloud_out[j] = tie_key_decode(ftk);
        } else {
            x1 = wslot[pb][src].x;
            y1 = wslot[pb][src].y;
            z1 = wslot[pb][src].z;
            if (t == 0) cloud_out[j] = tie_key_decode(ctk);
        }
    }
    if constexpr (CLUSTER) {
        // no CTA may exit while a peer could still write into its shared memory
        cg::this_cluster().sync();
    }
}

// Fallback for clouds beyond the register-resident capacity (n > 65536): running minima in a global scratch row
// per cloud, xyz streamed from L2 each round.  Same reduction machinery, one CTA of 1024 threads per cloud.
__global__ void __launch_bounds__(1024) fps_generic_kernel(int n, int m, const float *__restrict__ xyz,
                                                          float *__restrict__ temp, int *__restrict__ out) {
    constexpr int NW = 32;
    __shared__ Slot wslot[2][NW];
    const int cloud = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    float *td = temp + (size_t)cloud * n;
    int *o = out + (size_t)cloud * m;
    for (int k = t; k < n; k += 1024) td[k] = 1e38f;
    float x1 = __ldg(p + 0), y1 = __ldg(p + 1), z1 = __ldg(p + 2);
    if (t == 0) o[0] = 0;
    for (int j = 1; j < m; ++j) {
        const int pb = j & 1;
        unsigned bbits = 0u, btk = 0xffffffffu;
        float bx = 0.f, by = 0.f, bz = 0.f;
        for (int k = t; k < n; k += 1024) {
            const float x2 = __ldg(p + k * 3 + 0), y2 = __ldg(p + k * 3 + 1), z2 = __ldg(p + k * 3 + 2);
            const float d2 = fminf(dist2_fma(x2 - x1, y2 - y1, z2 - z1), td[k]);
            td[k] = d2;
            const unsigned bits = __float_as_uint(d2), tk = tie_key(k);
            if (bits > bbits || (bits == bbits && tk < btk)) {
                bbits = bits; btk = tk; bx = x2; by = y2; bz = z2;
            }
        }
        unsigned wbits, wtk;
        const bool win = warp_argmax(bbits, btk, wbits, wtk);
        if (lane == 0) { wslot[pb][warp].bits = wbits; wslot[pb][warp].tk = wtk; }
        if (win) { wslot[pb][warp].x = bx; wslot[pb][warp].y = by; wslot[pb][warp].z = bz; }
        __syncthreads();
        unsigned sbits = wslot[pb][lane].bits, stk = wslot[pb][lane].tk, cbits, ctk;
        warp_argmax(sbits, stk, cbits, ctk);
        int src = max(__ffs(__ballot_sync(FULL_MASK, sbits == cbits && stk == ctk)) - 1, 0);
        x1 = wslot[pb][src].x; y1 = wslot[pb][src].y; z1 = wslot[pb][src].z;
        if (t == 0) o[j] = tie_key_decode(ctk);
    }
}

template <int P, int THREADS>
int launch_fps(int b, int n, int m, int csize, int chunk, const float *xyz, int *out, cudaStream_t st) {
    const size_t smem = ((size_t)chunk * 3 + 8) * sizeof(float);
    cudaError_t e;
    if (csize == 1) {
        auto kern = fps_kernel<P, THREADS, false>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        kern<<<b, THREADS, smem, st>>>(n, m, chunk, xyz, out);
    } else {
        auto kern = fps_kernel<P, THREADS, true>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)b * csize);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)csize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, n, m, chunk, xyz, out);
        if (e != cudaSuccess) return (int)e;
    }
    return pasnl_launch_status();
}

constexpr int kMaxChunk = 8192;                          // points per CTA (1024 threads x 8)
constexpr int kMaxResident = kMaxChunk * kMaxCluster;    // 65536

}  // namespace

extern "C" size_t pasnl_fps_workspace_bytes(int b, int n, int m) {
    (void)m;
    if (b <= 0 || n <= kMaxResident) return 0;
    return (size_t)b * (size_t)n * sizeof(float);
}

extern "C" int pasnl_fps(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 1 && m >= 0);
    if (b == 0 || m == 0) return PASNL_OK;
    PASNL_CHECK_ARG(xyz != nullptr && out_idx != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    if (n > kMaxResident) {
        PASNL_CHECK_ARG(workspace != nullptr);
        fps_generic_kernel<<<b, 1024, 0, st>>>(n, m, xyz, (float *)workspace, out_idx);
        return pasnl_launch_status();
    }
    // Cluster size.  FPS is a latency chain of m rounds; a round costs (register sweep over the CTA's points) +
    // (intra-CTA reduction) [+ DSMEM exchange].  Up to 2048 points one CTA is fastest; above that the sweep dominates,
    // so the cloud is spread over a cluster (<= 8 CTAs, the portable maximum) of ~1280-point CTAs, growing the per-CTA
    // slice only when 8 CTAs are not enough.
    int csize = 1;
    if (n > 2048) {
        while (csize < kMaxCluster && (long)csize * 1280 < n) csize *= 2;
    }
    int chunk = (n + csize - 1) / csize;
    chunk = (chunk + 3) & ~3;  // keeps every slice 16-byte aligned relative to the cloud start
    // threads / points-per-thread for the per-CTA slice
#define FPS_CASE(P_, T_) return launch_fps<P_, T_>(b, n, m, csize, chunk, xyz, out_idx, st)
    if (chunk <= 128) FPS_CASE(1, 128);
    if (chunk <= 256) FPS_CASE(2, 128);
    if (chunk <= 512) FPS_CASE(4, 128);
    if (chunk <= 1024) FPS_CASE(4, 256);
    if (chunk <= 1280) FPS_CASE(5, 256);
    if (chunk <= 2048) FPS_CASE(4, 512);
    if (chunk <= 3072) FPS_CASE(6, 512);
    if (chunk <= 4096) FPS_CASE(4, 1024);
    if (chunk <= 6144) FPS_CASE(6, 1024);
    FPS_CASE(8, 1024);
#undef FPS_CASE
}
