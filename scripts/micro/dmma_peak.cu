// FP64 tensor (mma.sync m8n8k4 f64) vs FP64 FMA throughput on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma(double* out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
    for (int t = 0; t < 8; ++t) { c[t][0] = t; c[t][1] = -t; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    if (s == 123.456) out[0] = s;
}
__global__ void dfma(double* out, int iters, double a, double b)
{
    double x[16];
    for (int t = 0; t < 16; ++t) x[t] = threadIdx.x + t;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 16; ++t) x[t] = fma(x[t], a, b);
    }
    double s = 0;
    for (int t = 0; t < 16; ++t) s += x[t];
    if (s == 123.456) out[0] = s;
}
int main()
{
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dmma<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double(blocks) * threads / 32);
        printf("DMMA m8n8k4: %.2f ms  %.2f TFLOP/s\n", ms, flop / ms / 1e9);
        cudaEventRecord(e0); dfma<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        flop = 2.0 * 16.0 * iters * double(blocks) * threads;
        printf("DFMA       : %.2f ms  %.2f TFLOP/s\n", ms, flop / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
