// dense_glue.cu -- fused epilogue of the 1x1-conv layers that surround the point-set ops.
//
// The reference builds every such layer as conv -> bias_add -> batch_norm (train mode, eps 1e-3, biased variance) ->
// activation (utils/tf_util.py:120-185, 512-531): in TF1 -- and in a plain PyTorch mirror -- that is ~8 passes over
// the (B,S,K,C) activation forward and ~9 backward (bias add, statistics, transform, ReLU; ReLU', two BN reductions,
// BN transform, bias-gradient sum), all HBM bound.  Here the GEMM output y = x.W is consumed by
//   forward : one statistics pass (shifted sums, fp32 partials -> fp64 atomics) + one normalise/activate pass;
//   backward: one reduction pass (sum dz, sum dz*xhat) + one apply pass producing dy for the two GEMMs.
// The conv bias cancels exactly inside train-mode batch norm (mean(y+b) = mean(y)+b), so it is folded analytically:
// the output does not depend on it and its gradient is identically zero; the caller keeps the variable.
// Layout: y (rows, c) row-major, channel fastest (NHWC flattened).  All kernels are HBM bound:
//   stats 4*rows*c bytes; apply 8*rows*c; bwd reduce 8*rows*c; bwd apply 12*rows*c.
#include <cstdlib>

#include <cooperative_groups.h>

#include "common.cuh"

using namespace pasnl;
namespace cg = cooperative_groups;

namespace {

constexpr int kMaxThreads = 256;

struct Shape {
    int cvec;     // channel groups per row (c/4 when vectorised, else c)
    int vec;      // 4 or 1
    int rpb;      // rows handled concurrently by one CTA
    int threads;  // cvec * rpb  (<= 1024)
};

Shape make_shape(int c, bool aligned) {
    Shape s;
    s.vec = (c % 4 == 0 && aligned) ? 4 : 1;
    s.cvec = c / s.vec;
    s.rpb = s.cvec >= kMaxThreads ? 1 : kMaxThreads / s.cvec;
    s.threads = s.cvec * s.rpb;
    return s;
}

__device__ __forceinline__ float activate(float z, int act) {
    if (act == 1) return fmaxf(z, 0.f);
    if (act == 2) return z > 0.f ? z : 0.2f * z;   // tf.nn.leaky_relu default alpha
    return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
    if (act == 1) return z > 0.f ? 1.f : 0.f;
    if (act == 2) return z > 0.f ? 1.f : 0.2f;
    return 1.f;
}

extern __shared__ float red[];   // [2][threads][VEC] partial sums of the reduction phases

// ---------------------------------------------------------------- forward statistics
// ws[0..c) += sum(y - shift), ws[c..2c) += sum((y - shift)^2), shift = y[0, :]
template <int VEC>
__device__ __forceinline__ void bn_stats_body(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                double *__restrict__ ws) {
    const int t = threadIdx.x, cv = t % cvec, ry = t / cvec;
    const int col = cv * VEC;
    float shift[VEC], s1[VEC], s2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        shift[v] = __ldg(y + col + v);
        s1[v] = s2[v] = 0.f;
    }
    const long chunk = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
    constexpr int U = 4;   // rows in flight per thread: the pass is pure streaming, MLP is what matters
    for (long rb = r0 + ry; rb < r1; rb += (long)rpb * U) {
        float x[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + (long)u * rpb;
            if (r < r1) {
                if constexpr (VEC == 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(y + r * c + col));
                    x[u][0] = q.x; x[u][1] = q.y; x[u][2] = q.z; x[u][3] = q.w;
                } else {
                    x[u][0] = __ldg(y + r * c + col);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) x[u][v] = shift[v];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float d = x[u][v] - shift[v];
                s1[v] += d;
                s2[v] = fmaf(d, d, s2[v]);
            }
    }
    // reduce over the rpb row-lanes of the CTA, then one fp64 atomic per channel
    float *a = red, *b = red + blockDim.x * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        a[t * VEC + v] = s1[v];
        b[t * VEC + v] = s2[v];
    }
    __syncthreads();
    if (ry == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            double d1 = 0.0, d2 = 0.0;
            for (int k = 0; k < rpb; ++k) {
                d1 += (double)a[(k * cvec + cv) * VEC + v];
                d2 += (double)b[(k * cvec + cv) * VEC + v];
            }
            atomicAdd(ws + col + v, d1);
            atomicAdd(ws + c + col + v, d2);
        }
    }
}

// out = act((y - mean) * rstd * gamma + beta).  Same (row-lane, channel-group) thread layout as the statistics
// pass: a thread keeps its channels' parameters in registers and streams rows, so the only memory instructions in
// the loop are the data loads/stores themselves.
// The finalisation of the statistics (mean / biased variance from the sums, moving averages) is folded in: every thread
// derives its channels' mean and variance from the fp64 sums, CTA 0 also publishes them (saved for the backward) and
// updates the moving averages -- one launch less per layer.
template <int VEC>
__device__ __forceinline__ void bn_apply_body(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                const double *__restrict__ ws, int shifted, float *__restrict__ mean,
                                float *__restrict__ var, const float *__restrict__ bias, float *__restrict__ running_mean,
                                float *__restrict__ running_var, float decay, float vcorr, const float *__restrict__ gamma,
                                const float *__restrict__ beta, float eps, int act, float *__restrict__ out) {
    const int t = threadIdx.x, cv = t % cvec, ry = t / cvec;
    const int col = cv * VEC;
    float mu[VEC], sc[VEC], be[VEC];
    const double n = (double)rows;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int ch = col + v;
        const double m = ws[ch] / n;
        double vv = ws[c + ch] / n - m * m;
        if (vv < 0.0) vv = 0.0;
        mu[v] = (float)((shifted ? (double)__ldg(y + ch) : 0.0) + m);
        const float vf = (float)vv;
        sc[v] = rsqrtf(vf + eps) * __ldg(gamma + ch);
        be[v] = __ldg(beta + ch);
        if (blockIdx.x == 0 && ry == 0) {
            mean[ch] = mu[v];
            var[ch] = vf;
            // moving averages as tf.contrib.layers.batch_norm keeps them (assign_moving_average).  vcorr = rows/(rows-1)
            // when the TF layer takes the fused kernel (rank-2 / rank-4 inputs: the moving variance is fed the
            // Bessel-corrected batch variance), 1 for the rank-3 conv1d layers (biased variance)
            if (running_mean) running_mean[ch] = decay * running_mean[ch] + (1.f - decay) * (mu[v] + (bias ? bias[ch] : 0.f));
            if (running_var) running_var[ch] = decay * running_var[ch] + (1.f - decay) * vf * vcorr;
        }
    }
    const long chunk = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
    constexpr int U = 4;
    for (long rb = r0 + ry; rb < r1; rb += (long)rpb * U) {
        float x[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + (long)u * rpb;
            if (r < r1) {
                if constexpr (VEC == 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(y + r * c + col));
                    x[u][0] = q.x; x[u][1] = q.y; x[u][2] = q.z; x[u][3] = q.w;
                } else {
                    x[u][0] = __ldg(y + r * c + col);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + (long)u * rpb;
            if (r < r1) {
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) o[v] = activate(fmaf(x[u][v] - mu[v], sc[v], be[v]), act);
                if constexpr (VEC == 4)
                    *reinterpret_cast<float4 *>(out + r * c + col) = make_float4(o[0], o[1], o[2], o[3]);
                else
                    out[r * c + col] = o[0];
            }
        }
    }
}

// ---------------------------------------------------------------- backward
// ws[0..c) += sum dz, ws[c..2c) += sum dz*xhat,  dz = dout * act'(z), z recomputed exactly as in the forward
template <int VEC>
__device__ __forceinline__ void bn_bwd_reduce_body(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                     const float *__restrict__ dout, const float *__restrict__ mean,
                                     const float *__restrict__ var, const float *__restrict__ gamma,
                                     const float *__restrict__ beta, float eps, int act, double *__restrict__ ws) {
    const int t = threadIdx.x, cv = t % cvec, ry = t / cvec;
    const int col = cv * VEC;
    float mu[VEC], rstd[VEC], sc[VEC], be[VEC], s1[VEC], s2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        mu[v] = __ldg(mean + col + v);
        rstd[v] = rsqrtf(__ldg(var + col + v) + eps);
        sc[v] = rstd[v] * __ldg(gamma + col + v);
        be[v] = __ldg(beta + col + v);
        s1[v] = s2[v] = 0.f;
    }
    const long chunk = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
    constexpr int U = 4;   // 8 x 16 bytes in flight per thread: this pass only streams
    for (long rb = r0 + ry; rb < r1; rb += (long)rpb * U) {
        float x[U][VEC], g[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + (long)u * rpb;
            if (r < r1) {
                if constexpr (VEC == 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(y + r * c + col));
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(dout + r * c + col));
                    x[u][0] = q.x; x[u][1] = q.y; x[u][2] = q.z; x[u][3] = q.w;
                    g[u][0] = w.x; g[u][1] = w.y; g[u][2] = w.z; g[u][3] = w.w;
                } else {
                    x[u][0] = __ldg(y + r * c + col);
                    g[u][0] = __ldg(dout + r * c + col);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { x[u][v] = mu[v]; g[u][v] = 0.f; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float xc = x[u][v] - mu[v];
                const float z = fmaf(xc, sc[v], be[v]);
                const float dz = g[u][v] * act_grad(z, act);
                s1[v] += dz;
                s2[v] = fmaf(dz, xc * rstd[v], s2[v]);
            }
    }
    float *a = red, *b = red + blockDim.x * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        a[t * VEC + v] = s1[v];
        b[t * VEC + v] = s2[v];
    }
    __syncthreads();
    if (ry == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            double d1 = 0.0, d2 = 0.0;
            for (int k = 0; k < rpb; ++k) {
                d1 += (double)a[(k * cvec + cv) * VEC + v];
                d2 += (double)b[(k * cvec + cv) * VEC + v];
            }
            atomicAdd(ws + col + v, d1);
            atomicAdd(ws + c + col + v, d2);
        }
    }
}

// dy = gamma * rstd * (dz - mean(dz) - xhat * mean(dz*xhat)); block 0 also emits dgamma / dbeta
template <int VEC>
__device__ __forceinline__ void bn_bwd_apply_body(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                    const float *__restrict__ dout, const float *__restrict__ mean,
                                    const float *__restrict__ var, const float *__restrict__ gamma,
                                    const float *__restrict__ beta, float eps, int act,
                                    const double *__restrict__ ws, float *__restrict__ dy,
                                    float *__restrict__ dgamma, float *__restrict__ dbeta) {
    if (blockIdx.x == 0)
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            dbeta[ch] = (float)ws[ch];
            dgamma[ch] = (float)ws[c + ch];
        }
    const int t = threadIdx.x, cv = t % cvec, ry = t / cvec;
    const int col = cv * VEC;
    const float inv_n = (float)(1.0 / (double)rows);
    float mu[VEC], rstd[VEC], sc[VEC], be[VEC], m1[VEC], m2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        mu[v] = __ldg(mean + col + v);
        rstd[v] = rsqrtf(__ldg(var + col + v) + eps);
        sc[v] = rstd[v] * __ldg(gamma + col + v);
        be[v] = __ldg(beta + col + v);
        m1[v] = (float)ws[col + v] * inv_n;
        m2[v] = (float)ws[c + col + v] * inv_n;
    }
    const long chunk = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
    constexpr int U = 2;
    for (long rb = r0 + ry; rb < r1; rb += (long)rpb * U) {
        float x[U][VEC], d[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + (long)u * rpb;
            if (r < r1) {
                if constexpr (VEC == 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4 *>(y + r * c + col));
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(dout + r * c + col));
                    x[u][0] = q.x; x[u][1] = q.y; x[u][2] = q.z; x[u][3] = q.w;
                    d[u][0] = w.x; d[u][1] = w.y; d[u][2] = w.z; d[u][3] = w.w;
                } else {
                    x[u][0] = __ldg(y + r * c + col);
                    d[u][0] = __ldg(dout + r * c + col);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long r = rb + (long)u * rpb;
            if (r < r1) {
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float xc = x[u][v] - mu[v];
                    const float z = fmaf(xc, sc[v], be[v]);
                    const float dz = d[u][v] * act_grad(z, act);
                    o[v] = sc[v] * (dz - m1[v] - xc * rstd[v] * m2[v]);
                }
                if constexpr (VEC == 4)
                    *reinterpret_cast<float4 *>(dy + r * c + col) = make_float4(o[0], o[1], o[2], o[3]);
                else
                    dy[r * c + col] = o[0];
            }
        }
    }
}

// ---------------------------------------------------------------- kernels
template <int VEC>
__global__ void bn_stats_kernel(long rows, int c, int cvec, int rpb, const float *__restrict__ y, double *__restrict__ ws) {
    bn_stats_body<VEC>(rows, c, cvec, rpb, y, ws);
}
template <int VEC>
__global__ void bn_apply_kernel(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                const double *__restrict__ ws, int shifted, float *__restrict__ mean,
                                float *__restrict__ var, const float *__restrict__ bias, float *__restrict__ running_mean,
                                float *__restrict__ running_var, float decay, float vcorr, const float *__restrict__ gamma,
                                const float *__restrict__ beta, float eps, int act, float *__restrict__ out) {
    bn_apply_body<VEC>(rows, c, cvec, rpb, y, ws, shifted, mean, var, bias, running_mean, running_var, decay, vcorr, gamma, beta,
                       eps, act, out);
}
template <int VEC>
__global__ void bn_bwd_reduce_kernel(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                     const float *__restrict__ dout, const float *__restrict__ mean,
                                     const float *__restrict__ var, const float *__restrict__ gamma,
                                     const float *__restrict__ beta, float eps, int act, double *__restrict__ ws) {
    bn_bwd_reduce_body<VEC>(rows, c, cvec, rpb, y, dout, mean, var, gamma, beta, eps, act, ws);
}
template <int VEC>
__global__ void bn_bwd_apply_kernel(long rows, int c, int cvec, int rpb, const float *__restrict__ y,
                                    const float *__restrict__ dout, const float *__restrict__ mean,
                                    const float *__restrict__ var, const float *__restrict__ gamma,
                                    const float *__restrict__ beta, float eps, int act,
                                    const double *__restrict__ ws, float *__restrict__ dy,
                                    float *__restrict__ dgamma, float *__restrict__ dbeta) {
    bn_bwd_apply_body<VEC>(rows, c, cvec, rpb, y, dout, mean, var, gamma, beta, eps, act, ws, dy, dgamma, dbeta);
}

// Small layers (a few MB): the three dependent launches (clear, reduce, apply) cost more than the data movement, so the
// phases run inside ONE cooperative launch of at most one CTA per SM, separated by grid barriers; the second pass
// re-reads the CTA's own rows from L2.
template <int VEC>
__global__ void __launch_bounds__(kMaxThreads) bn_fwd_coop_kernel(
    long rows, int c, int cvec, int rpb, const float *__restrict__ y, double *ws, float *__restrict__ mean,
    float *__restrict__ var, const float *__restrict__ bias, float *__restrict__ running_mean,
    float *__restrict__ running_var, float decay, float vcorr, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
    int act, float *__restrict__ out) {
    cg::grid_group grid = cg::this_grid();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * c; i += gridDim.x * blockDim.x) ws[i] = 0.0;
    grid.sync();
    bn_stats_body<VEC>(rows, c, cvec, rpb, y, ws);
    grid.sync();
    bn_apply_body<VEC>(rows, c, cvec, rpb, y, ws, 1, mean, var, bias, running_mean, running_var, decay, vcorr, gamma, beta, eps,
                       act, out);
}
template <int VEC>
__global__ void __launch_bounds__(kMaxThreads) bn_bwd_coop_kernel(
    long rows, int c, int cvec, int rpb, const float *__restrict__ y, const float *__restrict__ dout,
    const float *__restrict__ mean, const float *__restrict__ var, const float *__restrict__ gamma,
    const float *__restrict__ beta, float eps, int act, double *ws, float *__restrict__ dy, float *__restrict__ dgamma,
    float *__restrict__ dbeta) {
    cg::grid_group grid = cg::this_grid();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * c; i += gridDim.x * blockDim.x) ws[i] = 0.0;
    grid.sync();
    bn_bwd_reduce_body<VEC>(rows, c, cvec, rpb, y, dout, mean, var, gamma, beta, eps, act, ws);
    grid.sync();
    bn_bwd_apply_body<VEC>(rows, c, cvec, rpb, y, dout, mean, var, gamma, beta, eps, act, ws, dy, dgamma, dbeta);
}

// layers up to this many elements take the single cooperative launch (PASNL_BN_COOP=0 disables it)
constexpr long kCoopMaxElems = 4L << 20;
bool coop_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("PASNL_BN_COOP");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
int coop_grid(long rows, int rpb) {
    long blocks = (rows + (long)rpb * 8 - 1) / ((long)rpb * 8);
    if (blocks > 148) blocks = 148;
    return (int)(blocks < 1 ? 1 : blocks);
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

int stream_grid(long work, int threads) {
    long blocks = (work + threads - 1) / threads;
    if (blocks > 148L * 32) blocks = 148L * 32;
    return (int)(blocks < 1 ? 1 : blocks);
}
int apply_grid(long rows, int rpb) {
    long blocks = (rows + (long)rpb * 8 - 1) / ((long)rpb * 8);     // >= 8 rows per thread-row
    if (blocks > 148L * 16) blocks = 148L * 16;
    return (int)(blocks < 1 ? 1 : blocks);
}
int reduce_grid(long rows, int rpb) {
    long blocks = (rows + (long)rpb * 8 - 1) / ((long)rpb * 8);     // >= 8 rows per thread-row (2 unrolled trips)
    static long cap = 0;
    if (cap == 0) {
        const char *e = getenv("PASNL_BN_RGRID");
        cap = e ? atol(e) : 2;   // measured (tools/sweep_bn.py): every CTA ends with 2c fp64 atomics on the same
        if (cap < 1) cap = 2;    // addresses; 2 CTAs per SM stream as fast as 8 and cut that contention (148 -> 130 us)
    }
    if (blocks > 148L * cap) blocks = 148L * cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" size_t pasnl_bn_workspace_bytes(int c) { return c > 0 ? sizeof(double) * 2 * (size_t)c : 0; }

extern "C" int pasnl_bn_act_fwd(long rows, int c, const float *y, const float *gamma, const float *beta, float eps,
                                int act, float *out, float *mean, float *var, const float *bias, float *running_mean,
                                float *running_var, float decay, int unbiased_moving_var, int stats_ready,
                                void *workspace, void *stream) {
    const float vcorr = (unbiased_moving_var && rows > 1) ? (float)((double)rows / (double)(rows - 1)) : 1.f;
    PASNL_CHECK_ARG(rows >= 1 && c >= 1 && act >= 0 && act <= 2 && c <= 4096);
    PASNL_CHECK_ARG(y && gamma && beta && out && mean && var && workspace);
    cudaStream_t st = (cudaStream_t)stream;
    double *ws = (double *)workspace;
    const Shape s = make_shape(c, aligned16(y) && aligned16(out));
    PASNL_CHECK_ARG(s.threads <= 1024);
    const size_t smem = sizeof(float) * 2 * (size_t)s.threads * s.vec;
    if (!stats_ready && coop_enabled() && rows * c <= kCoopMaxElems && s.threads <= kMaxThreads) {
        // small layer, statistics not provided: clear + statistics + normalise/activate in one cooperative launch
        long rows_ = rows;
        int c_ = c, cvec_ = s.cvec, rpb_ = s.rpb, act_ = act;
        float eps_ = eps, decay_ = decay, vcorr_ = vcorr;
        void *args[] = {&rows_, &c_, &cvec_, &rpb_, (void *)&y, &ws, &mean, &var, (void *)&bias, &running_mean,
                        &running_var, &decay_, &vcorr_, (void *)&gamma, (void *)&beta, &eps_, &act_, &out};
        const void *fn = s.vec == 4 ? (const void *)bn_fwd_coop_kernel<4> : (const void *)bn_fwd_coop_kernel<1>;
        cudaError_t ce = cudaLaunchCooperativeKernel(fn, dim3(coop_grid(rows, s.rpb)), dim3(s.threads), args, smem, st);
        if (ce != cudaSuccess) return (int)ce;
        return pasnl_launch_status();
    }
    if (!stats_ready) {
        cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)c, st);
        if (e != cudaSuccess) return (int)e;
    }
    const int rg = reduce_grid(rows, s.rpb);
    const int ag = apply_grid(rows, s.rpb);
    if (s.vec == 4) {
        if (!stats_ready) bn_stats_kernel<4><<<rg, s.threads, smem, st>>>(rows, c, s.cvec, s.rpb, y, ws);
        bn_apply_kernel<4><<<ag, s.threads, 0, st>>>(rows, c, s.cvec, s.rpb, y, ws, stats_ready ? 0 : 1, mean, var, bias,
                                                     running_mean, running_var, decay, vcorr, gamma, beta, eps, act, out);
    } else {
        if (!stats_ready) bn_stats_kernel<1><<<rg, s.threads, smem, st>>>(rows, c, s.cvec, s.rpb, y, ws);
        bn_apply_kernel<1><<<ag, s.threads, 0, st>>>(rows, c, s.cvec, s.rpb, y, ws, stats_ready ? 0 : 1, mean, var, bias,
                                                     running_mean, running_var, decay, vcorr, gamma, beta, eps, act, out);
    }
    return pasnl_launch_status();
}

extern "C" int pasnl_bn_act_bwd(long rows, int c, const float *y, const float *dout, const float *gamma,
                                const float *beta, const float *mean, const float *var, float eps, int act, float *dy,
                                float *dgamma, float *dbeta, int workspace_zeroed, void *workspace, void *stream) {
    PASNL_CHECK_ARG(rows >= 1 && c >= 1 && act >= 0 && act <= 2 && c <= 4096);
    PASNL_CHECK_ARG(y && dout && gamma && beta && mean && var && dy && dgamma && dbeta && workspace);
    cudaStream_t st = (cudaStream_t)stream;
    double *ws = (double *)workspace;
    const Shape s = make_shape(c, aligned16(y) && aligned16(dout) && aligned16(dy));
    PASNL_CHECK_ARG(s.threads <= 1024);
    const size_t smem = sizeof(float) * 2 * (size_t)s.threads * s.vec;
    if (coop_enabled() && rows * c <= kCoopMaxElems && s.threads <= kMaxThreads) {
        long rows_ = rows;
        int c_ = c, cvec_ = s.cvec, rpb_ = s.rpb, act_ = act;
        float eps_ = eps;
        void *args[] = {&rows_, &c_, &cvec_, &rpb_, (void *)&y, (void *)&dout, (void *)&mean, (void *)&var, (void *)&gamma,
                        (void *)&beta, &eps_, &act_, &ws, &dy, &dgamma, &dbeta};
        const void *fn = s.vec == 4 ? (const void *)bn_bwd_coop_kernel<4> : (const void *)bn_bwd_coop_kernel<1>;
        cudaError_t ce = cudaLaunchCooperativeKernel(fn, dim3(coop_grid(rows, s.rpb)), dim3(s.threads), args, smem, st);
        if (ce != cudaSuccess) return (int)ce;
        return pasnl_launch_status();
    }
    cudaError_t e = workspace_zeroed ? cudaSuccess : cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)c, st);
    if (e != cudaSuccess) return (int)e;
    const int rg = reduce_grid(rows, s.rpb);
    const int ag = apply_grid(rows, s.rpb);
    if (s.vec == 4) {
        bn_bwd_reduce_kernel<4><<<rg, s.threads, smem, st>>>(rows, c, s.cvec, s.rpb, y, dout, mean, var, gamma, beta, eps,
                                                            act, ws);
        bn_bwd_apply_kernel<4><<<ag, s.threads, 0, st>>>(rows, c, s.cvec, s.rpb, y, dout, mean, var, gamma, beta, eps, act,
                                                        ws, dy, dgamma, dbeta);
    } else {
        bn_bwd_reduce_kernel<1><<<rg, s.threads, smem, st>>>(rows, c, s.cvec, s.rpb, y, dout, mean, var, gamma, beta, eps,
                                                            act, ws);
        bn_bwd_apply_kernel<1><<<ag, s.threads, 0, st>>>(rows, c, s.cvec, s.rpb, y, dout, mean, var, gamma, beta, eps, act,
                                                        ws, dy, dgamma, dbeta);
    }
    return pasnl_launch_status();
}
