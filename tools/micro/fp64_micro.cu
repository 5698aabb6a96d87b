// Micro-benchmark (measurement aid): fp64 FMA latency and throughput on this GPU, one warp per
// SM sub-partition vs full occupancy.  nvcc -arch=sm_100a -O3 -o fp64_micro fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc)
{
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int blocks, int threads, const char *name)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<blocks, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<ILP><<<blocks, threads>>>(out, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double fma = (double)blocks * threads * iters * ILP;
    printf("%-28s ILP %d blocks %d thr %d: %.1f cycles per dependent step, %.2f TFMA/s (%.2f TFLOP/s)\n", name, ILP, blocks, threads,
           (double)h / iters, fma / (ms * 1e-3) / 1e12, 2 * fma / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<1>(1, 32, "1 warp, dependent chain");
    run<4>(1, 32, "1 warp, 4 chains");
    run<9>(1, 32, "1 warp, 9 chains");
    run<1>(1, 128, "4 warps (1/SMSP)");
    run<1>(1, 512, "16 warps");
    run<4>(1, 512, "16 warps, 4 chains");
    run<8>(148 * 4, 512, "full chip");
    return 0;
}
