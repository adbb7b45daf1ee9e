// grid_subsampling.cu -- voxel-grid barycentre subsampling on the GPU, sm_100a.
//
// Replaces utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106 (single-threaded C++ over an
// unordered_map<size_t, SampledData>): every point falls into the cell
//     (floor((p - origin) / dl)),   origin = floor(min_corner * (1 / dl)) * dl,   key = iX + NX*iY + NX*NY*iZ
// and every occupied cell yields the barycentre of its points, the mean of their features and, per label column, the
// most frequent label.  All arithmetic is fp32 in the reference's order: the cell sums are accumulated in ASCENDING
// POINT INDEX (the order the sequential loop visits them), so barycentres and feature means are bit-exact.
// Two things the reference leaves to its hash map's iteration order are pinned here: output cells are
// emitted in order of first appearance in the input, and a tie between label counts goes to the smallest label.
//
// HBM-bound hash / scan work: (1) min / max corner reduction; (2) cell key per point, open-addressing insert
// (atomicCAS on 64-bit keys), per-cell count and first point index; (3) prefix sum of "is the first point of its cell"
// -> cell rank; prefix sum of counts -> cell offsets; (4) scatter of point indices into per-cell lists; (5) one thread
// per cell: sort its (short) list, accumulate in index order, emit.
#include "common.cuh"

using namespace pasnl;

namespace gsub {

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr int kScanBlock = 1024;

__device__ __forceinline__ int float_ordered(float f) {          // monotone float -> int map for atomicMin / atomicMax
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void init_kernel(int *corner, unsigned long long *keys, long cap, int *cell_cnt, int *cell_first, int n,
                            int *num_out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3) corner[i] = 0x7FFFFFFF;
    else if (i < 6) corner[i] = (int)0x80000000;
    if (i == 0) *num_out = 0;
    if (i < cap) { keys[i] = kEmpty; cell_cnt[i] = 0; cell_first[i] = n; }
}

__global__ void minmax_kernel(int n, const float *__restrict__ pts, int *corner) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = __ldg(pts + 3 * i + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(FULL_MASK, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL_MASK, mx[d], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(corner + d, float_ordered(mn[d]));
            atomicMax(corner + 3 + d, float_ordered(mx[d]));
        }
    }
}

struct Grid {
    float ox, oy, oz;
    unsigned long long nx, ny;
};
// the reference's arithmetic, statement by statement (grid_subsampling.cpp:24-31): fp32 throughout
__device__ __forceinline__ Grid make_grid(const int *corner, float dl) {
    const float inv = 1 / dl;
    Grid g;
    g.ox = __fmul_rn(floorf(__fmul_rn(ordered_float(corner[0]), inv)), dl);
    g.oy = __fmul_rn(floorf(__fmul_rn(ordered_float(corner[1]), inv)), dl);
    g.oz = __fmul_rn(floorf(__fmul_rn(ordered_float(corner[2]), inv)), dl);
    g.nx = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(ordered_float(corner[3]), g.ox), dl)) + 1ull;
    g.ny = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(ordered_float(corner[4]), g.oy), dl)) + 1ull;
    return g;
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void insert_kernel(int n, const float *__restrict__ pts, float dl, const int *__restrict__ corner,
                              unsigned long long *keys, long cap, int *cell_cnt, int *cell_first, int *slot_of) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Grid g = make_grid(corner, dl);
    const unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i), g.ox), dl));
    const unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + 1), g.oy), dl));
    const unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(__ldg(pts + 3 * i + 2), g.oz), dl));
    const unsigned long long key = ix + g.nx * iy + g.nx * g.ny * iz;
    long slot = (long)(mix64(key) & (unsigned long long)(cap - 1));
    while (true) {
        const unsigned long long prev = atomicCAS(keys + slot, kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        slot = (slot + 1) & (cap - 1);
    }
    slot_of[i] = (int)slot;
    atomicAdd(cell_cnt + slot, 1);
    atomicMin(cell_first + slot, (int)i);
}

// flag[i] = 1 if point i is the first (lowest-index) point of its cell
__global__ void first_flag_kernel(int n, const int *__restrict__ slot_of, const int *__restrict__ cell_first, int *flag) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (cell_first[slot_of[i]] == (int)i) ? 1 : 0;
}

// ---- exclusive prefix sum of `n` ints, three launches: per-block scan + block totals, scan of totals, add back
__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(int n, const int *__restrict__ in, int *out, int *totals) {
    __shared__ int ws[32];
    const long i = (long)blockIdx.x * kScanBlock + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = ws[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL_MASK, w, o);
            if (threadIdx.x >= o) w += y;
        }
        ws[threadIdx.x] = w;
    }
    __syncthreads();
    const int base = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) out[i] = base + x - v;
    if (threadIdx.x == kScanBlock - 1) totals[blockIdx.x] = base + x;
}
__global__ void __launch_bounds__(kScanBlock) scan_totals_kernel(int nblocks, int *totals, int *grand) {
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += kScanBlock) {
        const int i = b0 + threadIdx.x;
        const int v = i < nblocks ? totals[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL_MASK, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = ws[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL_MASK, w, o);
                if (threadIdx.x >= o) w += y;
            }
            ws[threadIdx.x] = w;
        }
        __syncthreads();
        const int base = carry + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0);
        if (i < nblocks) totals[i] = base + x - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry = base + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && grand) *grand = carry;
}
__global__ void scan_add_kernel(int n, int *out, const int *__restrict__ totals) {
    const long i = (long)blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += totals[blockIdx.x];
}

// cell rank = number of earlier first-points; counts in rank order
__global__ void rank_kernel(int n, const int *__restrict__ flag, const int *__restrict__ rank_at, const int *__restrict__ slot_of,
                            const int *__restrict__ cell_cnt, int *cell_rank, int *vcnt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) {
        const int s = slot_of[i], r = rank_at[i];
        cell_rank[s] = r;
        vcnt[r] = cell_cnt[s];
    }
}
__global__ void fill_kernel(int n, const int *__restrict__ slot_of, const int *__restrict__ cell_rank,
                            const int *__restrict__ voff, int *cursor, int *list) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = cell_rank[slot_of[i]];
    list[voff[r] + atomicAdd(cursor + r, 1)] = (int)i;
}

// one thread per cell: ascending point order, then the reference's accumulation (grid_subsampling.h: SampledData)
__global__ void finalize_kernel(const int *__restrict__ num_cells, const int *__restrict__ voff, const int *__restrict__ vcnt,
                                int *list, const float *__restrict__ pts, int fdim, const float *__restrict__ feats, int ldim,
                                const int *__restrict__ classes, float *__restrict__ out_pts, float *__restrict__ out_feats,
                                int *__restrict__ out_classes) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *num_cells) return;
    int *l = list + voff[r];
    const int cnt = vcnt[r];
    for (int a = 1; a < cnt; ++a) {                       // insertion sort: cells hold a handful of points
        const int v = l[a];
        int b = a - 1;
        while (b >= 0 && l[b] > v) { l[b + 1] = l[b]; --b; }
        l[b + 1] = v;
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int a = 0; a < cnt; ++a) {
        const long i = l[a];
        sx = __fadd_rn(sx, __ldg(pts + 3 * i));
        sy = __fadd_rn(sy, __ldg(pts + 3 * i + 1));
        sz = __fadd_rn(sz, __ldg(pts + 3 * i + 2));
    }
    const float w = (float)(1.0 / (double)cnt);            // `point * (1.0 / count)`: the double is narrowed by operator*(PointXYZ, float)
    out_pts[3 * r] = __fmul_rn(sx, w);
    out_pts[3 * r + 1] = __fmul_rn(sy, w);
    out_pts[3 * r + 2] = __fmul_rn(sz, w);
    for (int f = 0; f < fdim; ++f) {
        float s = 0.f;
        for (int a = 0; a < cnt; ++a) s = __fadd_rn(s, __ldg(feats + (long)l[a] * fdim + f));
        out_feats[r * fdim + f] = __fdiv_rn(s, (float)cnt);
    }
    for (int c = 0; c < ldim; ++c) {                       // most frequent label; ties -> smallest label
        int best = 0, best_cnt = 0;
        for (int a = 0; a < cnt; ++a) {
            const int v = __ldg(classes + (long)l[a] * ldim + c);
            int k = 0;
            for (int b = 0; b < cnt; ++b) k += (__ldg(classes + (long)l[b] * ldim + c) == v) ? 1 : 0;
            if (k > best_cnt || (k == best_cnt && v < best)) { best = v; best_cnt = k; }
        }
        out_classes[r * ldim + c] = best;
    }
}

inline long capacity(int n) {
    long c = 64;
    while (c < 2L * n) c <<= 1;
    return c;
}
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct Layout {
    size_t corner, keys, cell_cnt, cell_first, cell_rank, slot_of, flag, rank_at, totals, vcnt, voff, cursor, list, total;
};
inline Layout layout(int n) {
    const long cap = capacity(n);
    const size_t nb = (size_t)((n + kScanBlock - 1) / kScanBlock) + 1;
    Layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += align256(bytes); return at; };
    L.corner = take(6 * 4);
    L.keys = take((size_t)cap * 8);
    L.cell_cnt = take((size_t)cap * 4);
    L.cell_first = take((size_t)cap * 4);
    L.cell_rank = take((size_t)cap * 4);
    L.slot_of = take((size_t)n * 4);
    L.flag = take((size_t)n * 4);
    L.rank_at = take((size_t)n * 4);
    L.totals = take(nb * 4);
    L.vcnt = take((size_t)n * 4);
    L.voff = take((size_t)n * 4);
    L.cursor = take((size_t)n * 4);
    L.list = take((size_t)n * 4);
    L.total = o;
    return L;
}

inline int exclusive_scan(int n, const int *in, int *out, int *totals, int *grand, cudaStream_t st) {
    const int nb = (n + kScanBlock - 1) / kScanBlock;
    scan_block_kernel<<<nb, kScanBlock, 0, st>>>(n, in, out, totals);
    scan_totals_kernel<<<1, kScanBlock, 0, st>>>(nb, totals, grand);
    scan_add_kernel<<<nb, kScanBlock, 0, st>>>(n, out, totals);
    return pasnl_launch_status();
}

}  // namespace gsub

extern "C" size_t pasnl_grid_subsample_workspace_bytes(int n) { return n > 0 ? gsub::layout(n).total : 0; }

extern "C" int pasnl_grid_subsample(int n, const float *points, int fdim, const float *features, int ldim, const int *classes,
                                    float sample_dl, void *workspace, int *num_out, float *sub_points, float *sub_features,
                                    int *sub_classes, void *stream) {
    PASNL_CHECK_ARG(n >= 0 && fdim >= 0 && ldim >= 0 && sample_dl > 0.f && num_out);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(num_out, 0, sizeof(int), st);
    PASNL_CHECK_ARG(points && workspace && sub_points && (fdim == 0 || (features && sub_features)) &&
                    (ldim == 0 || (classes && sub_classes)));
    using namespace gsub;
    const Layout L = layout(n);
    uint8_t *ws = (uint8_t *)workspace;
    int *corner = (int *)(ws + L.corner);
    unsigned long long *keys = (unsigned long long *)(ws + L.keys);
    int *cell_cnt = (int *)(ws + L.cell_cnt), *cell_first = (int *)(ws + L.cell_first), *cell_rank = (int *)(ws + L.cell_rank);
    int *slot_of = (int *)(ws + L.slot_of), *flag = (int *)(ws + L.flag), *rank_at = (int *)(ws + L.rank_at);
    int *totals = (int *)(ws + L.totals), *vcnt = (int *)(ws + L.vcnt), *voff = (int *)(ws + L.voff);
    int *cursor = (int *)(ws + L.cursor), *list = (int *)(ws + L.list);
    const long cap = capacity(n);
    const unsigned pb = (unsigned)((n + 255) / 256);
    init_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(corner, keys, cap, cell_cnt, cell_first, n, num_out);
    cudaError_t e = cudaMemsetAsync(ws + L.vcnt, 0, L.list - L.vcnt, st);      // vcnt, voff, cursor
    if (e != cudaSuccess) return (int)e;
    minmax_kernel<<<pb < 592 ? pb : 592, 256, 0, st>>>(n, points, corner);
    insert_kernel<<<pb, 256, 0, st>>>(n, points, sample_dl, corner, keys, cap, cell_cnt, cell_first, slot_of);
    first_flag_kernel<<<pb, 256, 0, st>>>(n, slot_of, cell_first, flag);
    int rc = exclusive_scan(n, flag, rank_at, totals, num_out, st);
    if (rc != 0) return rc;
    rank_kernel<<<pb, 256, 0, st>>>(n, flag, rank_at, slot_of, cell_cnt, cell_rank, vcnt);
    rc = exclusive_scan(n, vcnt, voff, totals, nullptr, st);
    if (rc != 0) return rc;
    fill_kernel<<<pb, 256, 0, st>>>(n, slot_of, cell_rank, voff, cursor, list);
    finalize_kernel<<<pb, 256, 0, st>>>(num_out, voff, vcnt, list, points, fdim, features, ldim, classes, sub_points,
                                        sub_features, sub_classes);
    return pasnl_launch_status();
}
