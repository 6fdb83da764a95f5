// Micro-benchmark: which sm_100a pipes execute packed 16-bit min/max, and can two forms co-issue?
// Decides how the median selection networks are emitted (rv_median_net.h, RV_MN / RV_MX).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_minmax ubench_minmax.cu && ./ubench_minmax
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int NP = 8;          // independent compare-exchange pairs per thread (ILP)
constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

struct OpVimnmx {   // VIMNMX.U16x2
    static __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) { uint32_t lo = __vminu2(a, b), hi = __vmaxu2(a, b); a = lo; b = hi; }
};
struct OpHmnmx {    // HMNMX2
    static __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) { __half2 x = u2h(a), y = u2h(b); a = h2u(__hmin2(x, y)); b = h2u(__hmax2(x, y)); }
};
struct OpHfma {     // fma pipe only: s = relu(b - a); max = a + s; min = b - s
    static __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b)
    {
        __half2 x = u2h(a), y = u2h(b);
        const __half2 one = __float2half2_rn(1.0f);
        __half2 s = __hfma2_relu(y, one, __hneg2(x));
        a = h2u(__hsub2(y, s)); b = h2u(__hadd2(x, s));
    }
};
struct OpFmnmx {    // fp32 min/max
    static __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) { float x = __uint_as_float(a), y = __uint_as_float(b); a = __float_as_uint(fminf(x, y)); b = __float_as_uint(fmaxf(x, y)); }
};
struct OpImnmx {    // 32-bit integer min/max
    static __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) { uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
};
struct OpVmin4 {    // __vminu4/__vmaxu4 (emulated on sm_100a)
    static __device__ __forceinline__ void ce(uint32_t &a, uint32_t &b) { uint32_t lo = __vminu4(a, b), hi = __vmaxu4(a, b); a = lo; b = hi; }
};

template <class A, class B, int NA>   // pairs [0, NA) use A, the rest use B
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[NP], b[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        uint32_t x = (threadIdx.x * 2654435761u + i * 40503u + seed) >> 7;
        a[i] = 0x64006400u | (x & 0x00ff00ffu);
        b[i] = 0x64006400u | ((x >> 8) & 0x00ff00ffu);
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < NA) A::ce(a[i], b[(i + 1) % NP]); else B::ce(a[i], b[(i + 1) % NP]);
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) r += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class A, class B, int NA>
void run(const char *name, uint32_t *out, int sms)
{
    dim3 grid(sms * 8), block(256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<A, B, NA><<<grid, block>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<A, B, NA><<<grid, block>>>(out, r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ce = 5.0 * grid.x * block.x * (double)ITERS * NP;
    printf("%-44s %8.3f ms  %8.1f G compare-exchange/s (thread-level, 2 lanes each)  err=%s\n", name, ms, ce / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    uint32_t *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    int s = p.multiProcessorCount;
    run<OpVimnmx, OpVimnmx, NP>("VIMNMX.U16x2 (vminu2/vmaxu2)", out, s);
    run<OpHmnmx, OpHmnmx, NP>("HMNMX2 (hmin2/hmax2)", out, s);
    run<OpVimnmx, OpHmnmx, NP / 2>("VIMNMX.U16x2 : HMNMX2 = 1:1", out, s);
    run<OpHfma, OpHfma, NP>("HFMA2.RELU+HADD2+HADD2 (fma pipe)", out, s);
    run<OpVimnmx, OpHfma, 5>("VIMNMX.U16x2 : HFMA2-CE = 5:3", out, s);
    run<OpVimnmx, OpHfma, 4>("VIMNMX.U16x2 : HFMA2-CE = 4:4", out, s);
    run<OpHmnmx, OpHfma, 5>("HMNMX2 : HFMA2-CE = 5:3", out, s);
    run<OpFmnmx, OpFmnmx, NP>("FMNMX (fp32, 1 lane)", out, s);
    run<OpImnmx, OpImnmx, NP>("IMNMX.U32 (1 lane)", out, s);
    run<OpVimnmx, OpFmnmx, NP / 2>("VIMNMX.U16x2 : FMNMX = 1:1", out, s);
    run<OpVmin4, OpVmin4, NP>("vminu4/vmaxu4 (emulated, 4 lanes)", out, s);
    return 0;
}
