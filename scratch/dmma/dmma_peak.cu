// FP64 tensor-core ceiling on this GPU: mma.sync m8n8k4 f64 from registers only, W warps per SM, C independent chains
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int C>
__global__ void k(double* out, int iters) {
    double c[C][2];
    for (int i = 0; i < C; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < C; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int C>
void run(int warps, int iters) {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<C><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e0);
    k<C><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 148.0 * warps * iters * C * 512.0;
    printf("warps/SM %2d chains %2d: %.3f ms  %.2f TFLOP/s  (%.2f clk per DMMA per SM at 1.965 GHz)\n", warps, C, ms, flops / ms / 1e9,
           ms * 1e-3 * 1.965e9 / ((double)warps * iters * C));
    cudaFree(out);
}
int main() {
    const int it = 20000;
    run<1>(4, it); run<1>(8, it); run<2>(8, it); run<4>(8, it); run<8>(8, it); run<8>(4, it); run<8>(16, it); run<24>(8, it);
    return 0;
}
