// neighbors.cu -- neighbour searches: ball query, exact k-NN, three-NN (sm_100a).
//
// Replaces query_ball_point_gpu (reference tf_ops/grouping/tf_grouping_g.cu:3-36: thread-per-query serial scan of
// all n support points from global memory, <<<b,256>>>), the nanoflann kd-tree k-NN the models call through
// tf.py_func on the HOST (utils/pointasnl_util.py:22-30,62 -> utils/nearest_neighbors/knn_.cxx:104-135) together with
// selection_sort_gpu (tf_grouping_g.cu:83-123), and threenn_cpu (tf_ops/3d_interpolation/tf_interpolate.cpp:60-103,
// one CPU thread).
//
// Common shape: a tiled pairwise-distance sweep.  grid = (query tiles, b); the support cloud is streamed through
// shared memory in tiles of kTile points staged with TMA bulk copies (cp.async.bulk + mbarrier, see common.cuh);
//   * ball query / k-NN: one WARP per query, lanes sweep 32 support points at a time (stride-3 float reads from
//     shared memory are bank-conflict free); ball query: ballot + popc ordered compaction ("first nsample in index
//     order"); k-NN: warp-select on 64-bit (distance, index) keys (queue + bitonic sort/merge in registers);
//   * three-NN: one THREAD per query, all lanes read the same support point (shared-memory broadcast).
// These sweeps are FP32/issue bound for n >~ 4k and latency bound below; compulsory HBM traffic is
// 12*b*(n+m) + outputs, the support tiles of all query tiles of a cloud are served from L2.
#include "common.cuh"

using namespace pasnl;

namespace {

constexpr int kWarps = 8;              // queries per CTA for the warp-per-query kernels
constexpr int kThreads = kWarps * 32;  // 256
constexpr int kTile = 2048;            // support points per shared-memory tile (24 KB)

struct TileStage {
    float buf[kTile * 3 + 8];
    uint64_t bar;
};

// ---------------------------------------------------------------- ball query
__global__ void __launch_bounds__(kThreads) ball_query_kernel(int n, int m, float radius, int nsample,
                                                              const float *__restrict__ xyz1,
                                                              const float *__restrict__ xyz2, int *__restrict__ idx,
                                                              int *__restrict__ pts_cnt) {
    __shared__ __align__(16) TileStage st;
    const int bi = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kWarps + warp;
    const bool valid = q < m;
    const float *sup = xyz1 + (size_t)bi * n * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float *qp = xyz2 + ((size_t)bi * m + q) * 3;
        qx = __ldg(qp);
        qy = __ldg(qp + 1);
        qz = __ldg(qp + 2);
    }
    int *orow = idx + ((size_t)bi * m + (valid ? q : 0)) * nsample;
    if (threadIdx.x == 0) {
        mbar_init(&st.bar, 1);
        fence_mbar_init();
    }
    int cnt = 0, first = 0;
    uint32_t parity = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int tile0 = 0; tile0 < n; tile0 += kTile) {
        // barrier: everybody finished reading the previous tile (and sees the mbarrier init on the first trip);
        // leave as soon as every query of the CTA is full.
        if (__syncthreads_and(!valid || cnt >= nsample)) break;
        const int tn = min(kTile, n - tile0);
        const float *pts = stage_floats(st.buf, sup + (size_t)tile0 * 3, tn * 3, &st.bar);
        stage_wait(&st.bar, parity);
        parity ^= 1u;
        if (valid && cnt < nsample) {
            for (int c0 = 0; c0 < tn; c0 += 32) {
                const int k = c0 + lane;
                bool hit = false;
                if (k < tn) {
                    // reference: (x2-x1) with x2 the query, x1 the support point (tf_grouping_g.cu:18-24)
                    const float d2 = dist2_fma(qx - pts[k * 3 + 0], qy - pts[k * 3 + 1], qz - pts[k * 3 + 2]);
                    hit = fmaxf(__fsqrt_rn(d2), 1e-20f) < radius;
                }
                const unsigned bal = __ballot_sync(FULL_MASK, hit);
                if (bal) {
                    const int pos = cnt + __popc(bal & lt_mask);
                    if (hit && pos < nsample) orow[pos] = tile0 + k;
                    if (cnt == 0) first = tile0 + c0 + __ffs(bal) - 1;
                    cnt += __popc(bal);
                    if (cnt >= nsample) break;
                }
            }
        }
    }
    if (valid) {
        cnt = min(cnt, nsample);
        if (lane == 0) pts_cnt[(size_t)bi * m + q] = cnt;
        // pad with the first hit (reference :26-29); rows without a hit are defined as 0
        for (int l = cnt + lane; l < nsample; l += 32) orow[l] = cnt > 0 ? first : 0;
    }
}

// ---------------------------------------------------------------- exact k-NN
// Warp-select.  Keys are 64-bit (distance bits << 32 | index): distances are non-negative, so unsigned order on the
// key IS the (distance, index) lexicographic order the contract asks for -- ties need no special handling.
// The current k best live in registers as a sorted list spread over the warp (position r*32 + lane, L = R*32 >= k
// entries, +inf padded).  Candidates below the k-th key are ballot-compacted into a per-warp shared-memory queue;
// whenever 32 are pending they are bitonic-sorted (15 shuffle stages) and bitonic-merged into the list (log2 L stages):
// ~200 warp instructions per 32 candidates instead of ~25 per candidate for one-at-a-time insertion.
constexpr unsigned long long kInfKey = 0x7f800000ffffffffull;   // (+inf, idx max)

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    const unsigned lo = __shfl_xor_sync(FULL_MASK, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(FULL_MASK, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}
// ascending bitonic sort of one key per lane
__device__ __forceinline__ unsigned long long warp_sort_asc(unsigned long long v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long o = shfl_xor_u64(v, j);
            const bool up = ((lane & k) == 0);            // ascending block
            const bool lower = ((lane & j) == 0);
            const bool take_min = (up == lower);
            v = take_min ? (v < o ? v : o) : (v > o ? v : o);
        }
    }
    return v;
}
// list (sorted ascending over positions r*32+lane) <- k smallest of (list U 32 sorted candidates c, one per lane)
template <int R>
__device__ __forceinline__ void merge_candidates(unsigned long long (&list)[R], unsigned long long c, int lane) {
    // element-wise min of the list's LAST 32 positions with the candidates reversed: the result is a bitonic
    // sequence containing the L smallest keys of the union (classic bitonic merge step, candidates padded with +inf)
    const unsigned long long crev = __shfl_sync(FULL_MASK, (unsigned)c, 31 - lane) |
                                    ((unsigned long long)__shfl_sync(FULL_MASK, (unsigned)(c >> 32), 31 - lane) << 32);
    list[R - 1] = list[R - 1] < crev ? list[R - 1] : crev;
    // bitonic merge network over L = R*32 positions: strides >= 32 pair registers of the same lane, strides < 32 shuffle
#pragma unroll
    for (int st = (R * 32) >> 1; st >= 32; st >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int pr = r ^ (st >> 5);
            if (pr > r) {
                const unsigned long long a = list[r], b2 = list[pr];
                list[r] = a < b2 ? a : b2;
                list[pr] = a < b2 ? b2 : a;
            }
        }
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned long long o = shfl_xor_u64(list[r], j);
            const bool lower = ((lane & j) == 0);
            list[r] = lower ? (list[r] < o ? list[r] : o) : (list[r] > o ? list[r] : o);
        }
    }
}

template <int R>
__global__ void __launch_bounds__(kThreads) knn_kernel(int n, int m, int k, const float *__restrict__ xyz1,
                                                       const float *__restrict__ xyz2, int *__restrict__ idx,
                                                       float *__restrict__ dist) {
    static_assert(R == 1 || R == 2 || R == 4, "list length must be a power of two");
    __shared__ __align__(16) TileStage st;
    __shared__ unsigned long long queue[kWarps][64];
    constexpr int U = 4;  // 32-point chunks in flight per lane
    const int bi_ = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kWarps + warp;
    const bool valid = q < m;
    const float *sup = xyz1 + (size_t)bi_ * n * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float *qp = xyz2 + ((size_t)bi_ * m + q) * 3;
        qx = __ldg(qp);
        qy = __ldg(qp + 1);
        qz = __ldg(qp + 2);
    }
    unsigned long long list[R];
#pragma unroll
    for (int r = 0; r < R; ++r) list[r] = kInfKey;
    unsigned long long tau = kInfKey;          // key at sorted position k-1: a candidate must be strictly below it
    int qn = 0;                                // queued candidates (warp-uniform)
    const int rk = (k - 1) >> 5, lk = (k - 1) & 31;
    unsigned long long *wq = queue[warp];
    const unsigned lt_mask = (1u << lane) - 1u;
    if (threadIdx.x == 0) {
        mbar_init(&st.bar, 1);
        fence_mbar_init();
    }
    uint32_t parity = 0;

    auto flush = [&](int take) {               // merge the first `take` (<= 32) queued keys into the list
        unsigned long long c = lane < take ? wq[lane] : kInfKey;
        c = warp_sort_asc(c, lane);
        merge_candidates<R>(list, c, lane);
        unsigned long long t = list[0];
#pragma unroll
        for (int r = 1; r < R; ++r) t = (r == rk) ? list[r] : t;
        const unsigned tlo = __shfl_sync(FULL_MASK, (unsigned)t, lk), thi = __shfl_sync(FULL_MASK, (unsigned)(t >> 32), lk);
        tau = ((unsigned long long)thi << 32) | tlo;
    };

    for (int tile0 = 0; tile0 < n; tile0 += kTile) {
        __syncthreads();
        const int tn = min(kTile, n - tile0);
        const float *pts = stage_floats(st.buf, sup + (size_t)tile0 * 3, tn * 3, &st.bar);
        stage_wait(&st.bar, parity);
        parity ^= 1u;
        if (!valid) continue;
        for (int c0 = 0; c0 < tn; c0 += 32 * U) {
            unsigned long long key[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = c0 + u * 32 + lane;
                key[u] = kInfKey;
                if (kk < tn) {
                    const float d = dist2_nofma(qx - pts[kk * 3 + 0], qy - pts[kk * 3 + 1], qz - pts[kk * 3 + 2]);
                    key[u] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(tile0 + kk);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool cand = key[u] < tau;
                const unsigned bal = __ballot_sync(FULL_MASK, cand);
                if (bal) {
                    if (cand) wq[qn + __popc(bal & lt_mask)] = key[u];
                    qn += __popc(bal);
                    __syncwarp();
                    if (qn >= 32) {
                        flush(32);
                        // slide the (< 32) remaining keys to the front; keys no longer below tau could be dropped,
                        // keeping them is harmless (they lose in the next merge)
                        const unsigned long long rest = (lane + 32 < qn) ? wq[lane + 32] : kInfKey;
                        __syncwarp();
                        wq[lane] = rest;
                        qn -= 32;
                        __syncwarp();
                    }
                }
            }
        }
    }
    if (valid) {
        if (qn > 0) flush(qn);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = r * 32 + lane;
            if (p < k) {
                const size_t o = ((size_t)bi_ * m + q) * k + p;
                idx[o] = (int)(unsigned)list[r];
                if (dist) dist[o] = __uint_as_float((unsigned)(list[r] >> 32));
            }
        }
    }
}

// ---------------------------------------------------------------- three-NN
__global__ void __launch_bounds__(kThreads) three_nn_kernel(int n, int m, const float *__restrict__ xyz1,
                                                            const float *__restrict__ xyz2, float *__restrict__ dist,
                                                            int *__restrict__ idx) {
    __shared__ __align__(16) TileStage st;
    const int bi = blockIdx.y;
    const int j = blockIdx.x * kThreads + threadIdx.x;
    const bool valid = j < n;
    const float *known = xyz2 + (size_t)bi * m * 3;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (valid) {
        const float *p = xyz1 + ((size_t)bi * n + j) * 3;
        x1 = __ldg(p);
        y1 = __ldg(p + 1);
        z1 = __ldg(p + 2);
    }
    const float inf = __int_as_float(0x7f800000);  // (float)1e40, the reference's initial "best"
    float b1 = inf, b2 = inf, b3 = inf;
    int i1 = 0, i2 = 0, i3 = 0;
    if (threadIdx.x == 0) {
        mbar_init(&st.bar, 1);
        fence_mbar_init();
    }
    uint32_t parity = 0;
    for (int tile0 = 0; tile0 < m; tile0 += kTile) {
        __syncthreads();
        const int tn = min(kTile, m - tile0);
        const float *pts = stage_floats(st.buf, known + (size_t)tile0 * 3, tn * 3, &st.bar);
        stage_wait(&st.bar, parity);
        parity ^= 1u;
#pragma unroll 4
        for (int k = 0; k < tn; ++k) {
            // reference: (x2-x1), x2 the known point (tf_interpolate.cpp:73); strict '<' cascade (:74-89)
            const float d = dist2_nofma(pts[k * 3 + 0] - x1, pts[k * 3 + 1] - y1, pts[k * 3 + 2] - z1);
            const int kk = tile0 + k;
            if (d < b1) {
                b3 = b2; i3 = i2;
                b2 = b1; i2 = i1;
                b1 = d; i1 = kk;
            } else if (d < b2) {
                b3 = b2; i3 = i2;
                b2 = d; i2 = kk;
            } else if (d < b3) {
                b3 = d; i3 = kk;
            }
        }
    }
    if (valid) {
        const size_t o = ((size_t)bi * n + j) * 3;
        dist[o] = b1; dist[o + 1] = b2; dist[o + 2] = b3;
        idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    }
}

}  // namespace

// ================================================================ C ABI
extern "C" int pasnl_ball_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                                int *idx, int *pts_cnt, void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0 && nsample >= 0);
    if ((long)b * m == 0) return PASNL_OK;
    PASNL_CHECK_ARG(xyz2 && pts_cnt && (nsample == 0 || idx) && (n == 0 || xyz1));
    PASNL_CHECK_ARG(b <= 65535);
    dim3 grid((unsigned)ceil_div_i(m, kWarps), (unsigned)b);
    ball_query_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
    return pasnl_launch_status();
}

extern "C" int pasnl_knn(int b, int n, int m, int k, const float *xyz1, const float *xyz2, int *idx, float *dist,
                         void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 1 && m >= 0 && k >= 1 && k <= n && k <= 128);
    if ((long)b * m == 0) return PASNL_OK;
    PASNL_CHECK_ARG(xyz1 && xyz2 && idx && b <= 65535);
    dim3 grid((unsigned)ceil_div_i(m, kWarps), (unsigned)b);
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 32) knn_kernel<1><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist);
    else if (k <= 64) knn_kernel<2><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist);
    else knn_kernel<4><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist);
    return pasnl_launch_status();
}

extern "C" int pasnl_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                              void *stream) {
    PASNL_CHECK_ARG(b >= 0 && n >= 0 && m >= 0);
    if ((long)b * n == 0) return PASNL_OK;
    PASNL_CHECK_ARG(xyz1 && dist && idx && (m == 0 || xyz2) && b <= 65535);
    dim3 grid((unsigned)ceil_div_i(n, kThreads), (unsigned)b);
    three_nn_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, m, xyz1, xyz2, dist, idx);
    return pasnl_launch_status();
}
