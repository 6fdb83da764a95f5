/*
 * rv_common.cuh -- shared by all kernels of the chain: frame / tile geometry, small helpers, TMA + mbarrier wrappers.
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cuda_fp16.h>

#include <type_traits>

namespace rv {

struct Geo {
    int H, W;          // frame size
    int grid;          // tiles per side
    int tw, th;        // tile size of the (REFLECT_101-padded) plane, clahe.cpp semantics (A.3)
    float inv_tw, inv_th;
};

// Shared-memory index checks for the debug build (make librv_b200_dbg.so: -DRV_DEBUG_BOUNDS).  compute-sanitizer is closed on the
// pool this code is developed on, so the hand-derived shared-memory bounds of k_chain / k_luma_hist / k_build_lut* are asserted
// in-kernel instead: a violation prints the site and traps (the launch fails with an error, it never corrupts silently).  The
// production build compiles the checks away.
#ifdef RV_DEBUG_BOUNDS
#define RV_CHECK_IDX(i, n, what)                                                                              \
    do {                                                                                                      \
        if ((long long)(i) < 0 || (long long)(i) >= (long long)(n)) {                                         \
            printf("RV_DEBUG_BOUNDS: %s index %lld outside [0,%lld) at %s:%d block (%d,%d,%d) thread %d\n", what, (long long)(i), \
                   (long long)(n), __FILE__, __LINE__, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);       \
            __trap();                                                                                         \
        }                                                                                                     \
    } while (0)
#else
#define RV_CHECK_IDX(i, n, what) ((void)0)
#endif

__device__ __forceinline__ int sat8(int v) { return __vimin_s32_relu(v, 255); }      // max(min(v, 255), 0): one VIMNMX.RELU

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) + mbarrier helpers for the box staging of k_chain
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1 << 22)) __trap();          // a lost TMA must fail loudly, never hang the GPU
    }
}
// 3-D tiled load: box -> dense shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no barrier): warms the tile a later CTA will stage
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

}  // namespace rv
