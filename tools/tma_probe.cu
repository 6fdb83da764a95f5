// Stand-alone probe of the k_chain TMA staging (3-D u32 tensor map, 96 x BOX_H x 1 box, negative start coords).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int rows, uint32_t *out, int *status)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(rows * 384) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t done = 0;
    int spin = 0;
    for (; !done && spin < (1 << 20); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) { status[0] = done; status[1] = spin; }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * 96; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(smem)[i];
}

typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int W = 384, H = 216, N = 2, ROWS = 34;
    const size_t pitch = 3 * W, fs = pitch * H;
    std::vector<uint8_t> h(fs * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint32_t *out; cudaMalloc(&out, ROWS * 384); int *st; cudaMalloc(&st, 8);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    EncFn enc = (EncFn)p;
    CUtensorMap map;
    cuuint64_t dims[3] = {3 * W / 4, H, N};
    cuuint64_t strides[2] = {pitch, fs};
    cuuint32_t box[3] = {96, ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 384);
    int cc[3] = {argc > 1 ? atoi(argv[1]) : 0, argc > 2 ? atoi(argv[2]) : 0, argc > 3 ? atoi(argv[3]) : 0};
    int coords[][3] = {{cc[0], cc[1], cc[2]}};
    for (auto &c : coords) {
        cudaMemset(out, 0xEE, ROWS * 384);
        probe<<<1, 128, ROWS * 384>>>(map, c[0], c[1], c[2], ROWS, out, st);
        e = cudaDeviceSynchronize();
        int hs[2]; cudaMemcpy(hs, st, 8, cudaMemcpyDeviceToHost);
        std::vector<uint32_t> ho(ROWS * 96); cudaMemcpy(ho.data(), out, ROWS * 384, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int ry = 0; ry < ROWS; ++ry)
            for (int wx = 0; wx < 96; ++wx) {
                int gy = c[1] + ry, gx = c[0] + wx;
                uint32_t want = 0;
                if (gy >= 0 && gy < H && gx >= 0 && gx < 3 * W / 4) memcpy(&want, &h[(size_t)c[2] * fs + gy * pitch + 4 * gx], 4);
                if (ho[ry * 96 + wx] != want) ++bad;
            }
        printf("coords (%d,%d,%d): sync=%s done=%d spins=%d mismatches=%ld\n", c[0], c[1], c[2], cudaGetErrorString(e), hs[0], hs[1], bad);
    }
    return 0;
}
