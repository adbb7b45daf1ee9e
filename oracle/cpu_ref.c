/*
 * oracle/cpu_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-file restatement of the reference (yanx27/PointASNL) point-set ops on the hot
 * path.  It exists only so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs can check / time the algorithm on host cores.  The product path
 * (pointasnl_b200/) never imports, links or calls anything in this directory.
 *
 * Parity status: the reference ships NO golden vectors or tests for these ops,
 * so the restatement is pinned against the reference itself, compiled unmodified into oracle/_ref/
 * by oracle/build_ref.sh:
 *   - knn / three_nn / three_interpolate(+grad): pinned HERE (CPU) against ref_knn.so / ref_interp.so,
 *     fixtures in tests/golden/ (tests/test_oracle_golden.py).
 *   - fps / gather / ball query / group(+grad): the reference has GPU kernels only; pinned on the GPU
 *     box against ref_sampling.so / ref_grouping.so (tests/test_gpu_parity.py) and through the golden
 *     fixtures those runs produced (tests/golden/gpu_ref_*.npz).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).  -ffp-contract=off matters:
 * FMA use is written out explicitly with fmaf() where the reference's GPU build contracts, and is
 * forbidden where the reference runs on x86 (no FMA in the baseline x86-64 ISA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* number of OpenMP threads used over the batch dimension (1 = the faithful single-thread baseline) */
static int g_threads = 1;
void oracle_set_threads(int t) { g_threads = t < 1 ? 1 : t; }
int oracle_get_threads(void) { return g_threads; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* squared distance exactly as nvcc contracts the reference expression
 *   (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1)
 * on the GPU: FMUL(dy*dy); FFMA(dx,dx,.); FFMA(dz,dz,.)  (SASS probe). */
static inline float d2_gpu(float x1, float y1, float z1, float x2, float y2, float z2) {
    float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}
/* squared distance as the x86 builds evaluate it: ((dx*dx + dy*dy) + dz*dz), no FMA
 * (tf_interpolate.cpp:73; nanoflann.hpp L2_Adaptor::evalMetric tail loop :341-344). */
static inline float d2_x86(float x1, float y1, float z1, float x2, float y2, float z2) {
    float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return (dx * dx + dy * dy) + dz * dz;
}

/* ------------------------------------------------------------------------------------------------
 * Farthest point sampling.  Follows tf_ops/sampling/tf_sampling_g.cu:105-170
 * (farthestpointsamplingKernel, launched <<<32,512>>> at :203-205).
 *   idx[0] = 0 (:114-116); temp init 1e38 (:118); per round d2 = min(d, temp[k]) (:144),
 *   thread t scans k = t, t+512, ... keeping the first strict maximum (:146-149), then a pairwise tree
 *   that keeps the LEFT operand on ties (:153-163).  Net tie rule: largest d2, then lowest (k mod 512),
 *   then lowest k.  This closed form is cross-checked against oracle_fps_sim512 (a literal simulation of
 *   the 512 threads and the tree) in tests/test_oracle.py.
 * ---------------------------------------------------------------------------------------------- */
void oracle_fps(int b, int n, int m, const float *xyz, int *idx) {
    if (m <= 0) return; /* :106 */
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i) {
        const float *p = xyz + (size_t)i * n * 3;
        int *o = idx + (size_t)i * m;
        float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        for (int k = 0; k < n; ++k) temp[k] = 1e38f;
        int old = 0;
        o[0] = 0;
        for (int j = 1; j < m; ++j) {
            float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            float best = -1.0f;
            int besti = 0, bestt = 0;
            for (int k = 0; k < n; ++k) {
                float d = d2_gpu(x1, y1, z1, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                int t = k & 511;
                if (d2 > best || (d2 == best && (t < bestt || (t == bestt && k < besti)))) {
                    best = d2;
                    besti = k;
                    bestt = t;
                }
            }
            old = besti;
            o[j] = old;
        }
        free(temp);
    }
}

/* Literal simulation of the 512-thread block + tree reduce of tf_sampling_g.cu:123-168.
 * Slow; used only to validate the closed-form tie rule of oracle_fps. */
void oracle_fps_sim512(int b, int n, int m, const float *xyz, int *idx) {
    if (m <= 0) return;
    enum { BS = 512 };
    for (int i = 0; i < b; ++i) {
        const float *p = xyz + (size_t)i * n * 3;
        int *o = idx + (size_t)i * m;
        float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        float dists[BS];
        int dists_i[BS];
        for (int k = 0; k < n; ++k) temp[k] = 1e38f;
        int old = 0;
        o[0] = 0;
        for (int j = 1; j < m; ++j) {
            float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int t = 0; t < BS; ++t) {
                int besti = 0;
                float best = -1;
                for (int k = t; k < n; k += BS) {
                    float td = temp[k];
                    float d = d2_gpu(x1, y1, z1, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                    float d2 = fminf(d, td);
                    if (d2 != td) temp[k] = d2;
                    if (d2 > best) {
                        best = d2;
                        besti = k;
                    }
                }
                dists[t] = best;
                dists_i[t] = besti;
            }
            for (int u = 0; (1 << u) < BS; ++u) {
                for (int t = 0; t < (BS >> (u + 1)); ++t) {
                    int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (dists[i1] < dists[i2]) {
                        dists[i1] = dists[i2];
                        dists_i[i1] = dists_i[i2];
                    }
                }
            }
            old = dists_i[0];
            o[j] = old;
        }
        free(temp);
    }
}

/* gather_point: tf_sampling_g.cu:172-181.  out[b,j,:] = inp[b,idx[b,j],:]  (3 floats) */
void oracle_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int l = 0; l < 3; ++l) out[((size_t)i * m + j) * 3 + l] = inp[((size_t)i * n + a) * 3 + l];
        }
}
/* gather_point_grad: tf_sampling_g.cu:183-192 (atomicAdd scatter; caller memset at tf_sampling.cpp:174).
 * The oracle zero-fills and accumulates sequentially in index order (fp32). */
void oracle_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int l = 0; l < 3; ++l) inp_g[((size_t)i * n + a) * 3 + l] += out_g[((size_t)i * m + j) * 3 + l];
        }
}

/* ------------------------------------------------------------------------------------------------
 * query_ball_point: tf_ops/grouping/tf_grouping_g.cu:3-36.
 *   scan support points in index order; stop once cnt == nsample (:16, tested BEFORE point k);
 *   predicate max(sqrtf(d2),1e-20f) < radius (:24-25); on the first hit all nsample slots are filled
 *   with k (:26-29); pts_cnt always written (:34).
 *   Rows with no hit: the reference leaves idx untouched (uninitialised TF output); here and in the
 *   product they are defined as 0 (documented deviation that no in-range query can observe).
 * ---------------------------------------------------------------------------------------------- */
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                       int *idx, int *pts_cnt) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i) {
        const float *s = xyz1 + (size_t)i * n * 3;
        const float *q = xyz2 + (size_t)i * m * 3;
        int *oi = idx + (size_t)i * m * nsample;
        int *oc = pts_cnt + (size_t)i * m;
        for (int j = 0; j < m; ++j) {
            int cnt = 0;
            for (int l = 0; l < nsample; ++l) oi[(size_t)j * nsample + l] = 0;
            float x2 = q[j * 3 + 0], y2 = q[j * 3 + 1], z2 = q[j * 3 + 2];
            for (int k = 0; k < n; ++k) {
                if (cnt == nsample) break;
                float x1 = s[k * 3 + 0], y1 = s[k * 3 + 1], z1 = s[k * 3 + 2];
                /* reference writes (x2-x1): query minus support */
                float d = fmaxf(sqrtf(d2_gpu(x1, y1, z1, x2, y2, z2)), 1e-20f);
                if (d < radius) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) oi[(size_t)j * nsample + l] = k;
                    oi[(size_t)j * nsample + cnt] = k;
                    cnt += 1;
                }
            }
            oc[j] = cnt;
        }
    }
}

/* group_point: tf_grouping_g.cu:40-57.  out[b,j,k,:] = points[b,idx[b,j,k],:] */
void oracle_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = idx[((size_t)i * m + j) * nsample + k];
                memcpy(out + (((size_t)i * m + j) * nsample + k) * c, points + ((size_t)i * n + ii) * c,
                       sizeof(float) * (size_t)c);
            }
}
/* group_point_grad: tf_grouping_g.cu:61-78 (atomicAdd; caller memset tf_grouping.cpp:204). */
void oracle_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                             float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = idx[((size_t)i * m + j) * nsample + k];
                const float *g = grad_out + (((size_t)i * m + j) * nsample + k) * c;
                float *d = grad_points + ((size_t)i * n + ii) * c;
                for (int l = 0; l < c; ++l) d[l] += g[l];
            }
}

/* ------------------------------------------------------------------------------------------------
 * k nearest neighbours, the contract of knn_query (utils/pointasnl_util.py:22-30 ->
 * utils/nearest_neighbors/knn_.cxx:104-135 -> nanoflann 1.2.3) and of knn_point / select_top_k
 * (tf_ops/grouping/tf_grouping.py:48-73 -> selection_sort_gpu tf_grouping_g.cu:83-123).
 *   distance = ((dx*dx + dy*dy) + dz*dz) in fp32 without FMA (nanoflann.hpp:341-344; exact search,
 *   eps = 0, knn_.cxx:127); result ascending by distance (KNNResultSet::addPoint, nanoflann.hpp:115-139).
 *   Ties: nanoflann keeps tree-visit order; selection_sort finds the earliest minimum (strict '<',
 *   tf_grouping_g.cu:110) but then SWAPS it into place, which reorders later equal elements.  Neither order is a
 *   function of the inputs a caller could rely on, so the oracle (and the product) define ties by ascending index;
 *   that equals both references whenever no two candidates of a query are equidistant, and always yields the same
 *   distances position by position.
 * ---------------------------------------------------------------------------------------------- */
void oracle_knn(int b, int n, int m, int k, const float *xyz1, const float *xyz2, int *idx, float *dist) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i) {
        const float *s = xyz1 + (size_t)i * n * 3;
        const float *q = xyz2 + (size_t)i * m * 3;
        float *bd = (float *)malloc(sizeof(float) * (size_t)k);
        int *bi = (int *)malloc(sizeof(int) * (size_t)k);
        for (int j = 0; j < m; ++j) {
            int cnt = 0;
            float qx = q[j * 3 + 0], qy = q[j * 3 + 1], qz = q[j * 3 + 2];
            for (int t = 0; t < n; ++t) {
                float d = d2_x86(s[t * 3 + 0], s[t * 3 + 1], s[t * 3 + 2], qx, qy, qz);
                if (cnt == k && !(d < bd[k - 1])) continue;
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && bd[pos - 1] > d) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = t;
                if (cnt < k) ++cnt;
            }
            for (int t = 0; t < k; ++t) {
                idx[((size_t)i * m + j) * k + t] = t < cnt ? bi[t] : 0;
                if (dist) dist[((size_t)i * m + j) * k + t] = t < cnt ? bd[t] : INFINITY;
            }
        }
        free(bd);
        free(bi);
    }
}

/* selection_sort (select_top_k): tf_grouping_g.cu:83-123.  Full (b,m,n) outputs; the first k entries of
 * each row are the k smallest in ascending order, the rest is the permuted remainder. */
void oracle_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *src = dist + ((size_t)i * m + j) * n;
            float *pd = out + ((size_t)i * m + j) * n;
            int *pi = outi + ((size_t)i * m + j) * n;
            for (int s = 0; s < n; ++s) {
                pd[s] = src[s];
                pi[s] = s;
            }
            for (int s = 0; s < k; ++s) {
                int mn = s;
                for (int t = s + 1; t < n; ++t)
                    if (pd[t] < pd[mn]) mn = t;
                if (mn != s) {
                    float tmp = pd[mn];
                    pd[mn] = pd[s];
                    pd[s] = tmp;
                    int ti = pi[mn];
                    pi[mn] = pi[s];
                    pi[s] = ti;
                }
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * three_nn: tf_ops/3d_interpolation/tf_interpolate.cpp:60-103.  Three smallest SQUARED distances
 * (the sqrt at :72 is commented out), strict '<' cascade so the earliest index wins ties; d evaluated
 * in float then widened to double for the compares (:73); best init 1e40 (-> +inf as float), idx init 0.
 * ---------------------------------------------------------------------------------------------- */
void oracle_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i) {
        const float *u = xyz1 + (size_t)i * n * 3;
        const float *kn = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; ++j) {
            float x1 = u[j * 3 + 0], y1 = u[j * 3 + 1], z1 = u[j * 3 + 2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                double d = d2_x86(x1, y1, z1, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; i3 = i2;
                    best2 = best1; i2 = i1;
                    best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2;
                    best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            size_t o = ((size_t)i * n + j) * 3;
            dist[o] = (float)best1; idx[o] = i1;
            dist[o + 1] = (float)best2; idx[o + 1] = i2;
            dist[o + 2] = (float)best3; idx[o + 2] = i3;
        }
    }
}

/* three_interpolate: tf_interpolate.cpp:107-127.  out = p1*w1 + p2*w2 + p3*w3, x86 evaluation
 * order ((p1*w1 + p2*w2) + p3*w3), no FMA. */
void oracle_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                              float *out) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            size_t o = ((size_t)i * n + j) * 3;
            float w1 = weight[o], w2 = weight[o + 1], w3 = weight[o + 2];
            const float *p1 = points + ((size_t)i * m + idx[o]) * c;
            const float *p2 = points + ((size_t)i * m + idx[o + 1]) * c;
            const float *p3 = points + ((size_t)i * m + idx[o + 2]) * c;
            float *d = out + ((size_t)i * n + j) * c;
            for (int l = 0; l < c; ++l) d[l] = (p1[l] * w1 + p2[l] * w2) + p3[l] * w3;
        }
}
/* three_interpolate_grad: tf_interpolate.cpp:131-153 (caller memset :258).  Sequential '+=' in j order. */
void oracle_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                   const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * m * c);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            size_t o = ((size_t)i * n + j) * 3;
            const float *g = grad_out + ((size_t)i * n + j) * c;
            for (int t = 0; t < 3; ++t) {
                float w = weight[o + t];
                float *d = grad_points + ((size_t)i * m + idx[o + t]) * c;
                for (int l = 0; l < c; ++l) d[l] += g[l] * w;
            }
        }
}


/* ---------------------------------------------------------------------------------------------- grid subsampling
 * utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106 with SampledData of
 * grid_subsampling.h:12-88: one pass over the points in index order, per cell count / fp32 point sum / fp32 feature
 * sums / label histograms, then barycentre = sum * (float)(1.0 / count), feature = sum / (float)count, label = most
 * frequent.  fp32 arithmetic statement by statement (:24-31 origin and grid size, :52-55 cell index).
 * Pinned where the reference is implementation-defined (std::unordered_map iteration order): cells are emitted in
 * order of first appearance, label ties go to the smallest label.  Returns the number of cells. */
static int cmp_key_idx(const void *a, const void *b) {
    const unsigned long long *x = (const unsigned long long *)a, *y = (const unsigned long long *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1] ? 1 : 0);
}
static int cmp_first(const void *a, const void *b) {
    const long *x = (const long *)a, *y = (const long *)b;
    return x[0] < y[0] ? -1 : (x[0] > y[0] ? 1 : 0);
}
int oracle_grid_subsample_ex(int n, const float *pts, int fdim, const float *feats, int ldim, const int *classes, float dl,
                             float *out_pts, float *out_feats, int *out_classes, int tie_largest);
int oracle_grid_subsample(int n, const float *pts, int fdim, const float *feats, int ldim, const int *classes, float dl,
                          float *out_pts, float *out_feats, int *out_classes) {
    return oracle_grid_subsample_ex(n, pts, fdim, feats, ldim, classes, dl, out_pts, out_feats, out_classes, 0);
}
/* tie_largest != 0: label ties go to the LARGEST label instead -- used by the tests to tell a tie (where the
 * reference's answer is implementation-defined) from a real disagreement */
int oracle_grid_subsample_ex(int n, const float *pts, int fdim, const float *feats, int ldim, const int *classes, float dl,
                             float *out_pts, float *out_feats, int *out_classes, int tie_largest) {
    if (n <= 0) return 0;
    float mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            if (pts[3 * i + d] < mn[d]) mn[d] = pts[3 * i + d];
            if (pts[3 * i + d] > mx[d]) mx[d] = pts[3 * i + d];
        }
    const float inv = 1 / dl;
    float org[3];
    for (int d = 0; d < 3; ++d) org[d] = floorf(mn[d] * inv) * dl;
    const unsigned long long nx = (unsigned long long)floorf((mx[0] - org[0]) / dl) + 1;
    const unsigned long long ny = (unsigned long long)floorf((mx[1] - org[1]) / dl) + 1;
    unsigned long long *ki = (unsigned long long *)malloc(sizeof(unsigned long long) * 2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const unsigned long long ix = (unsigned long long)floorf((pts[3 * i] - org[0]) / dl);
        const unsigned long long iy = (unsigned long long)floorf((pts[3 * i + 1] - org[1]) / dl);
        const unsigned long long iz = (unsigned long long)floorf((pts[3 * i + 2] - org[2]) / dl);
        ki[2 * i] = ix + nx * iy + nx * ny * iz;
        ki[2 * i + 1] = (unsigned long long)i;
    }
    qsort(ki, (size_t)n, 2 * sizeof(unsigned long long), cmp_key_idx);      /* cells contiguous, points ascending inside */
    long *cells = (long *)malloc(sizeof(long) * 3 * (size_t)n);             /* (first index, start, count) */
    int nc = 0;
    for (int a = 0; a < n;) {
        int b = a;
        while (b < n && ki[2 * b] == ki[2 * a]) ++b;
        cells[3 * nc] = (long)ki[2 * a + 1];
        cells[3 * nc + 1] = a;
        cells[3 * nc + 2] = b - a;
        ++nc;
        a = b;
    }
    qsort(cells, (size_t)nc, 3 * sizeof(long), cmp_first);                  /* order of first appearance */
    for (int c = 0; c < nc; ++c) {
        const long a0 = cells[3 * c + 1], cnt = cells[3 * c + 2];
        float s[3] = {0.f, 0.f, 0.f};
        for (long a = a0; a < a0 + cnt; ++a) {
            const long i = (long)ki[2 * a + 1];
            s[0] += pts[3 * i]; s[1] += pts[3 * i + 1]; s[2] += pts[3 * i + 2];
        }
        const float w = (float)(1.0 / (double)cnt);
        out_pts[3 * c] = s[0] * w; out_pts[3 * c + 1] = s[1] * w; out_pts[3 * c + 2] = s[2] * w;
        for (int f = 0; f < fdim; ++f) {
            float t = 0.f;
            for (long a = a0; a < a0 + cnt; ++a) t += feats[(long)ki[2 * a + 1] * fdim + f];
            out_feats[(long)c * fdim + f] = t / (float)cnt;
        }
        for (int l = 0; l < ldim; ++l) {
            int best = 0, bestc = 0;
            for (long a = a0; a < a0 + cnt; ++a) {
                const int v = classes[(long)ki[2 * a + 1] * ldim + l];
                int k = 0;
                for (long b = a0; b < a0 + cnt; ++b) k += classes[(long)ki[2 * b + 1] * ldim + l] == v;
                if (k > bestc || (k == bestc && (tie_largest ? v > best : v < best))) { best = v; bestc = k; }
            }
            out_classes[(long)c * ldim + l] = best;
        }
    }
    free(ki);
    free(cells);
    return nc;
}


/* ---------------------------------------------------------------------------------------------- prob_sample
 * tf_sampling_g.cu:7-104 (cumsumKernel + binarysearchKernel).  The index returned depends on the rounding of the fp32
 * running sums, so the summation tree is restated exactly: 8192-element blocks; inside a block groups of four
 * (v1, v1+v2, v3+(v1+v2), (v4+v3)+(v2+v1)); the group totals go through an up-sweep / down-sweep scan (:44-63); element =
 * (prefix in group + scanned total of the previous group) + carry; the carry between blocks is compensated (:73-76). */
void oracle_cumsum(int b, int n, const float *inp, float *out) {
    enum { BLK = 8192 };
    float *pre = (float *)malloc(sizeof(float) * BLK), *tot = (float *)malloc(sizeof(float) * (BLK / 4));
    for (int i = 0; i < b; ++i) {
        const float *src = inp + (size_t)i * n;
        float *dst = out + (size_t)i * n;
        float carry = 0.f, comp = 0.f;
        for (int j = 0; j < n; j += BLK) {
            const int cnt = n - j < BLK ? n - j : BLK;
            const int padded = (cnt + 3) & ~3, groups = padded >> 2;
            for (int g = 0; g < groups; ++g) {
                const int k = g * 4;
                if (k + 3 < cnt) {
                    const float v1 = src[j + k], v2 = src[j + k + 1], v3 = src[j + k + 2], v4 = src[j + k + 3];
                    const float s12 = v2 + v1, s34 = v4 + v3;
                    pre[k] = v1; pre[k + 1] = s12; pre[k + 2] = v3 + s12; pre[k + 3] = s34 + s12;
                    tot[g] = pre[k + 3];
                } else {
                    float v = 0.f;
                    for (int e = k; e < cnt; ++e) { v += src[j + e]; pre[e] = v; }
                    for (int e = cnt; e < padded; ++e) pre[e] = v;
                    tot[g] = v;
                }
            }
            int u = 0;
            for (; (2 << u) <= groups; ++u)
                for (int q = 0; q < (groups >> (u + 1)); ++q)
                    tot[(((q << 1) + 2) << u) - 1] += tot[(((q << 1) + 1) << u) - 1];
            for (--u; u >= 0; --u)
                for (int q = 0; q < ((groups - (1 << u)) >> (u + 1)); ++q)
                    tot[(((q << 1) + 3) << u) - 1] += tot[(((q << 1) + 2) << u) - 1];
            for (int e = 0; e < cnt; ++e) {
                const int g = e >> 2;
                const float in_block = g > 0 ? pre[e] + tot[g - 1] : pre[e];
                dst[j + e] = in_block + carry;
            }
            const float t = tot[groups - 1] + comp;
            const float r2 = carry + t;
            comp = t - (r2 - carry);
            carry = r2;
        }
    }
    free(pre);
    free(tot);
}
void oracle_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out) {
    oracle_cumsum(b, n, inp_p, temp);
    int base = 1;
    while (base < n) base <<= 1;
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *row = temp + (size_t)i * n;
            const float q = inp_r[(size_t)i * m + j] * row[n - 1];
            int r = n - 1;
            for (int k = base; k >= 1; k >>= 1)
                if (r >= k && row[r - k] >= q) r -= k;
            out[(size_t)i * m + j] = r;
        }
}
