// Speed of light of the median phase of k_chain: the generated 5x5 two-row selection network (rv_median5x2_net, same
// compare-exchange macros and ALU/FMA mix as the product kernel) with its inputs read from conflict-free per-thread shared memory and
// nothing else -- no staging, no colour work, no byte stores, no barriers -- at the product's occupancy (256 threads x 3 CTAs per SM, 80 registers).  One call of
// the network is one "task" of k_chain's phase 2 (2 x 6 output words = 24 output bytes of one channel plane); a 1080p
// frame is 259,200 tasks.  The result bounds k_chain from below and says how much of its time is pure network.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRV_CHAIN_MIN_CTAS=3 -DRV_FMA_NUM=1 -DRV_FMA_DEN=2 \
//        -I../road-vision-system_b200/csrc -o ubench_median_sol ubench_median_sol.cu && ./ubench_median_sol
#include <stdio.h>
#include "rv_kernels.cuh"

constexpr int ITERS = 512;

template <int K>
__global__ void __launch_bounds__(256, 3) k_sol(uint32_t *out, uint32_t seed)
{
    constexpr int M = (K == 5) ? RV_MEDIAN5X2_M : RV_MEDIAN3X2_M;
    constexpr int NC = M + K - 1, NR = K + 1;
    // every thread owns NC x NR words of shared memory (stride 256 words: conflict-free), read with one 32-bit load per
    // input (the product reads the same words with 64-bit loads from planes shared between threads)
    extern __shared__ uint32_t sm[];
    volatile uint32_t *mine = sm + threadIdx.x;      // volatile: every call re-reads all its inputs (nothing is hoisted out of the loop)
#pragma unroll
    for (int i = 0; i < NC * NR; ++i) {
        const uint32_t x = (threadIdx.x * 2654435761u + i * 40503u + seed + blockIdx.x) >> 3;
        mine[i * 256] = RV_PLANE_BIAS | (x & 0x00ff00ffu);
    }
    for (int it = 0; it < ITERS; ++it) {
        uint32_t v[NC][NR];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int r = 0; r < NR; ++r) v[c][r] = mine[(c * NR + r) * 256];
        uint32_t o[2][M];
        if constexpr (K == 5) rv_median5x2_net(v, o); else rv_median3x2_net(v, o);
        // feed the outputs back as the outer rows of the first M columns: keeps every call live and data dependent
#pragma unroll
        for (int j = 0; j < M; ++j) { mine[(j * NR) * 256] = o[0][j]; mine[(j * NR + NR - 1) * 256] = o[1][j]; }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NC * NR; ++i) r += mine[i * 256];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int K>
void run(const char *name, uint32_t *out, int sms, double tasks_per_frame)
{
    dim3 grid(sms * 3 * 4), block(256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    constexpr int SM_BYTES = ((K == 5 ? RV_MEDIAN5X2_M : RV_MEDIAN3X2_M) + K - 1) * (K + 1) * 256 * 4;
    cudaFuncSetAttribute(k_sol<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BYTES);
    k_sol<K><<<grid, block, SM_BYTES>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k_sol<K><<<grid, block, SM_BYTES>>>(out, r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tasks = 5.0 * grid.x * block.x * (double)ITERS;
    const double rate = tasks / (ms * 1e-3);
    printf("%-28s %8.3f ms  %8.2f G tasks/s  -> %7.2f us per 1080p frame (%.0f tasks), i.e. at most %8.0f frames/s for this phase alone  err=%s\n",
           name, ms, rate / 1e9, tasks_per_frame / rate * 1e6, tasks_per_frame, rate / tasks_per_frame, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz; compare-exchange mix RV_FMA %d/%d\n", p.name, p.multiProcessorCount, p.clockRate / 1000, RV_FMA_NUM, RV_FMA_DEN);
    uint32_t *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 12 * 256 * 4);
    // tasks per 1080p frame: 3 channels x (1920 / 6) groups x (1080 / 4) row pairs of two packed rows
    run<5>("5x5 two-row network (k5)", out, p.multiProcessorCount, 3.0 * (1920 / 6) * (1080 / 4));
    run<3>("3x3 two-row network (k3)", out, p.multiProcessorCount, 3.0 * (1920 / 6) * (1080 / 4));
    return 0;
}
