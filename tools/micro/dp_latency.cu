// Micro-benchmark: fp64 dependent-chain latency and pipe throughput on this GPU (one warp / many warps).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double *out, int iters, double a, double b)
{
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = __dadd_rn(x[k], b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / iters;
}
template <int ILP>
__global__ void chain_shfl(double *out, int iters, double b)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __dadd_rn(__shfl_up_sync(0xffffffffu, x, 1), b);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0) / iters;
}
int main()
{
    double *d; cudaMalloc(&d, 1 << 24);
    double h;
    const int it = 4096;
#define RUN(K, blocks, threads, label) chain<K><<<blocks, threads>>>(d, it, 1.0, 1e-9); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); \
    printf("%-44s %7.2f cycles per loop iteration (%d DADD per thread)\n", label, h, K);
    RUN(1, 1, 32, "1 warp, 1 dependent chain")
    RUN(2, 1, 32, "1 warp, 2 independent chains")
    RUN(4, 1, 32, "1 warp, 4 independent chains")
    RUN(8, 1, 32, "1 warp, 8 independent chains")
    RUN(1, 1, 128, "4 warps (1/SMSP), 1 chain each")
    RUN(1, 1, 512, "16 warps (4/SMSP), 1 chain each")
    RUN(2, 1, 512, "16 warps (4/SMSP), 2 chains each")
    RUN(4, 1, 512, "16 warps, 4 chains each")
    RUN(8, 1, 1024, "32 warps, 8 chains each (pipe throughput)")
    chain_shfl<1><<<1, 32>>>(d, it, 1e-9); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.2f cycles (64-bit shuffle + DADD dependent)\n", "1 warp, shfl+dadd chain", h);
    return 0;
}
