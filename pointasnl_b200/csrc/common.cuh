// common.cuh -- shared device helpers for the sm_100a kernels of libpasnl_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pasnl.h"

#define PASNL_CHECK_ARG(cond) \
    do {                      \
        if (!(cond)) return PASNL_EINVAL; \
    } while (0)

// Launch epilogue: the ABI reports the launch status, never synchronises.
static inline int pasnl_launch_status() { return (int)cudaGetLastError(); }

static inline int ceil_div_i(long a, long b) { return (int)((a + b - 1) / b); }

#define FULL_MASK 0xffffffffu

namespace pasnl {

// ------------------------------------------------------------------------------------------------
// Distance arithmetic.  Both forms are spelled with explicit intrinsics so that neither -fmad nor the
// optimiser can change the rounding sequence the parity contract fixes.
// ------------------------------------------------------------------------------------------------
// The reference's GPU kernels (FPS, ball query) as nvcc contracts them: FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
__device__ __forceinline__ float dist2_fma(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// The reference's x86 code (nanoflann kNN, three_nn): ((dx*dx + dy*dy) + dz*dz), no FMA.
__device__ __forceinline__ float dist2_nofma(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// ------------------------------------------------------------------------------------------------
// TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared with an mbarrier, used to stage point
// tiles.  Bulk copies need 16-byte aligned addresses and sizes, but a cloud inside a (b,n,3) tensor starts at
// b*n*12 bytes, which is only 4-byte aligned in general.  stage_floats() therefore places the data in shared
// memory with the SAME 16-byte phase as the global source: element t of the source lands at smem[phase + t],
// phase = (addr % 16) / 4.  The aligned interior goes through one bulk copy, the <= 3 leading and <= 3
// trailing floats through ordinary loads.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// phase (in floats) the staged copy of `g` will have inside a 16-byte aligned shared buffer
__device__ __forceinline__ int stage_phase(const float *g) { return (int)(((uintptr_t)g & 15u) >> 2); }

// Cooperative: every thread of the CTA calls this with identical arguments.  `buf` is 16-byte aligned and holds
// at least nfloats + 3 floats.  Returns the pointer at which element 0 of the source is found.  The caller must
// have executed a CTA barrier since anybody last read `buf`, and must call stage_wait() before reading it.
// `bar` must have been initialised with count 1; `parity` is the phase the caller tracks (flips per use).
__device__ __forceinline__ const float *stage_floats(float *buf, const float *g, int nfloats, uint64_t *bar) {
    const int phase = stage_phase(g);
    float *dst = buf + phase;
    int head = (4 - phase) & 3;
    if (head > nfloats) head = nfloats;
    const int body = (nfloats - head) & ~3;
    const int tail0 = head + body;
    if (threadIdx.x == 0) {
        if (body > 0) {
            mbar_expect_tx(bar, (uint32_t)body * 4u);
            // one bulk copy may carry at most 2^20-1 transaction bytes per mbarrier phase; tiles here are far smaller
            bulk_g2s(dst + head, g + head, (uint32_t)body * 4u, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    if ((int)threadIdx.x < head) dst[threadIdx.x] = g[threadIdx.x];
    if ((int)threadIdx.x >= 32 && (int)threadIdx.x - 32 < nfloats - tail0)
        dst[tail0 + threadIdx.x - 32] = g[tail0 + threadIdx.x - 32];
    return dst;
}
// Wait for the bulk part, then barrier so that the head/tail stores of other threads are visible too.
__device__ __forceinline__ void stage_wait(uint64_t *bar, uint32_t parity) {
    mbar_wait(bar, parity);
    __syncthreads();
}

}  // namespace pasnl
