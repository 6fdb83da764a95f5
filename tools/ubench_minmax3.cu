// Micro-benchmark: issue rate of the three-input packed min/max VIMNMX3.U16x2 on sm_100a (ptxas fuses
// __vminu2(__vminu2(a,b),c) into it) against the two-input VIMNMX.U16x2, alone and mixed with FMA-pipe work.
// Decides whether tools/gen_median_net.py should fuse min(min(a,b),c) / max(max(a,b),c) chains.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_minmax3 ubench_minmax3.cu && ./ubench_minmax3
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int NP = 8;
constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

// MODE 0: 2 x VIMNMX (2-input) per pair-iteration; 1: 2 x VIMNMX3; 2: 2 x VIMNMX3 + 2 x HFMA2; 3: 2 x VIMNMX + 2 x HFMA2;
// 4: 2 x IADD3 ; 5: VIMNMX3 + IADD3 ; 6: 2 x HFMA2 only
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[NP], b[NP], c[NP], h[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        uint32_t x = (threadIdx.x * 2654435761u + i * 40503u + seed) >> 5;
        a[i] = 0x64006400u | (x & 0x00ff00ffu);
        b[i] = 0x64006400u | ((x >> 8) & 0x00ff00ffu);
        c[i] = 0x64006400u | ((x >> 3) & 0x00ff00ffu);
        h[i] = 0x3c003c00u | ((x >> 4) & 0x00030003u);
    }
    const __half2 one = __float2half2_rn(1.0f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int n = (i + 1) % NP;
            if (MODE == 0 || MODE == 3) {
                a[i] = __vminu2(a[i], b[n]);
                b[i] = __vmaxu2(b[i], c[n]);
            }
            if (MODE == 1 || MODE == 2) {
                a[i] = __vminu2(__vminu2(a[i], b[n]), c[i]);
                b[i] = __vmaxu2(__vmaxu2(b[i], c[n]), a[n]);
            }
            if (MODE == 2 || MODE == 3 || MODE == 6) {
                h[i] = h2u(__hfma2(u2h(h[i]), one, u2h(h[n])));
                c[i] = h2u(__hfma2(u2h(c[i]), one, u2h(h[i])));
            }
            if (MODE == 4) {
                a[i] = a[i] + b[n] + c[i];
                b[i] = b[i] + c[n] + a[n];
            }
            if (MODE == 5) {
                a[i] = __vminu2(__vminu2(a[i], b[n]), c[i]);
                b[i] = b[i] + c[n] + a[n];
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) r += a[i] + b[i] + c[i] + h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int instr_per_pair, uint32_t *out, int sms)
{
    dim3 grid(sms * 8), block(256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(out, r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ins = 5.0 * grid.x * block.x * (double)ITERS * NP * instr_per_pair;
    printf("%-44s %8.3f ms  %8.1f G thread-instr/s = %5.1f lanes/clk/SM  err=%s\n", name, ms, ins / ms / 1e6,
           ins / ms / 1e6 * 1e9 / (sms * 1.965e9), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    uint32_t *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    int s = p.multiProcessorCount;
    run<0>("2 x VIMNMX.U16x2", 2, out, s);
    run<1>("2 x VIMNMX3.U16x2", 2, out, s);
    run<6>("2 x HFMA2", 2, out, s);
    run<3>("2 x VIMNMX.U16x2 + 2 x HFMA2", 4, out, s);
    run<2>("2 x VIMNMX3.U16x2 + 2 x HFMA2", 4, out, s);
    run<4>("2 x IADD3", 2, out, s);
    run<5>("VIMNMX3.U16x2 + IADD3", 2, out, s);
    return 0;
}
