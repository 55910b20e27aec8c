// How many warps does it take to keep the FP64 tensor path (mma.sync m8n8k4 f64, SASS DMMA) busy?
// One block per SM, W warps per block, every warp issues U independent accumulators round robin.
// Prints cycles per DMMA per SM sub-partition (4 per SM): 16 = the pipe is saturated.
#include <cstdio>
#include <cuda_runtime.h>
template <int U> __global__ void dmma(double* out, int iters, long long* cycles)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[U][2];
#pragma unroll
    for (int t = 0; t < U; ++t) { c[t][0] = t; c[t][1] = -t; }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < U; ++t) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int t = 0; t < U; ++t) s += c[t][0] + c[t][1];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int U> void run(int warps, double* d, long long* dc)
{
    const int iters = 4000;
    dmma<U><<<148, warps * 32>>>(d, iters, dc);
    cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    const double per_subpartition = double(cyc) / (double(iters) * U * warps / 4.0);
    printf("warps/SM %2d  independent accumulators %d : %.1f cycles per DMMA per sub-partition, %.1f cycles per DMMA per warp\n",
           warps, U, per_subpartition, double(cyc) / (double(iters) * U));
}
int main()
{
    double* d; cudaMalloc(&d, 8);
    long long* dc; cudaMalloc(&dc, 8);
    for (int warps : {4, 8, 16, 32}) {
        run<1>(warps, d, dc);
        run<2>(warps, d, dc);
        run<4>(warps, d, dc);
        run<8>(warps, d, dc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
