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
//     shared memory are bank-conflict free), ballot + popc give the ordered compaction the "first nsample in index
//     order" / "ascending (d, index)" contracts need;
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
// Warp-distributed sorted list: position p = r*32 + lane holds the p-th nearest so far.
template <int R>
__device__ __forceinline__ void knn_insert(float (&bd)[R], int (&bi)[R], float dc, int ic, int lane) {
    int pos = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) pos += __popc(__ballot_sync(FULL_MASK, bd[r] <= dc));
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        float ud = __shfl_up_sync(FULL_MASK, bd[r], 1);
        int ui = __shfl_up_sync(FULL_MASK, bi[r], 1);
        if (r > 0) {
            const float cd = __shfl_sync(FULL_MASK, bd[r - 1], 31);
            const int ci = __shfl_sync(FULL_MASK, bi[r - 1], 31);
            if (lane == 0) {
                ud = cd;
                ui = ci;
            }
        }
        const int p = r * 32 + lane;
        if (p == pos) {
            bd[r] = dc;
            bi[r] = ic;
        } else if (p > pos) {
            bd[r] = ud;
            bi[r] = ui;
        }
    }
}

template <int R>
__global__ void __launch_bounds__(kThreads) knn_kernel(int n, int m, int k, const float *__restrict__ xyz1,
                                                       const float *__restrict__ xyz2, int *__restrict__ idx,
                                                       float *__restrict__ dist) {
    __shared__ __align__(16) TileStage st;
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
    float bd[R];
    int bi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        bd[r] = __int_as_float(0x7f800000);
        bi[r] = 0;
    }
    float tau = __int_as_float(0x7f800000);  // distance at sorted position k-1
    int nfill = 0;
    const int rk = (k - 1) >> 5, lk = (k - 1) & 31;
    if (threadIdx.x == 0) {
        mbar_init(&st.bar, 1);
        fence_mbar_init();
    }
    uint32_t parity = 0;
    for (int tile0 = 0; tile0 < n; tile0 += kTile) {
        __syncthreads();
        const int tn = min(kTile, n - tile0);
        const float *pts = stage_floats(st.buf, sup + (size_t)tile0 * 3, tn * 3, &st.bar);
        stage_wait(&st.bar, parity);
        parity ^= 1u;
        if (!valid) continue;
        for (int c0 = 0; c0 < tn; c0 += 32 * U) {
            float d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = c0 + u * 32 + lane;
                d[u] = __int_as_float(0x7fc00000);  // NaN: never a candidate
                if (kk < tn) d[u] = dist2_nofma(qx - pts[kk * 3 + 0], qy - pts[kk * 3 + 1], qz - pts[kk * 3 + 2]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in_range = c0 + u * 32 + lane < tn;
                unsigned bal = __ballot_sync(FULL_MASK, in_range && (d[u] < tau || nfill < k));
                while (bal) {  // ascending lane == ascending support index: later equal distances go behind
                    const int src = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const float dc = __shfl_sync(FULL_MASK, d[u], src);
                    if (dc < tau || nfill < k) {
                        knn_insert<R>(bd, bi, dc, tile0 + c0 + u * 32 + src, lane);
                        nfill = min(nfill + 1, k);
                        float t = bd[0];
#pragma unroll
                        for (int r = 1; r < R; ++r) t = (r == rk) ? bd[r] : t;
                        tau = __shfl_sync(FULL_MASK, t, lk);
                    }
                }
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = r * 32 + lane;
            if (p < k) {
                const size_t o = ((size_t)bi_ * m + q) * k + p;
                idx[o] = bi[r];
                if (dist) dist[o] = bd[r];
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
    switch ((k + 31) / 32) {
        case 1: knn_kernel<1><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist); break;
        case 2: knn_kernel<2><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist); break;
        case 3: knn_kernel<3><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist); break;
        default: knn_kernel<4><<<grid, kThreads, 0, st>>>(n, m, k, xyz1, xyz2, idx, dist); break;
    }
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
