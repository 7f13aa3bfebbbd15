// Microbenchmark: FP64 mma.sync m8n8k4 (DMMA) throughput on sm_100a, alone and concurrently with DFMA chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NMMA, int NFMA>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double c[8][2];
    double x[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; x[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NMMA; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < NFMA; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + x[i];
    if (s == 1234.5) out[0] = s;
}
template <int NMMA, int NFMA> void run(const char* name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = sms * blocks_per_sm;
    k<NMMA, NFMA><<<blocks, 256>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); k<NMMA, NFMA><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = blocks * 8.0;
    const double mma = warps * iters * 4.0 * NMMA, fma_w = warps * iters * 4.0 * NFMA;
    printf("%-28s %7.3f ms  DMMA %.2f TFLOP/s (%.3f per clk per SM)  DFMA %.2f TFLOP/s  err=%s\n", name, ms, mma * 512 / ms * 1e-9,
           mma / (ms * 1e-3 * 1.965e9 * sms), fma_w * 64 / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<8, 0>("DMMA x8 chains, 1 CTA/SM", 1);
    run<8, 0>("DMMA x8 chains, 4 CTA/SM", 4);
    run<0, 8>("DFMA x8 chains, 4 CTA/SM", 4);
    run<8, 8>("DMMA x8 + DFMA x8, 4 CTA/SM", 4);
    run<4, 8>("DMMA x4 + DFMA x8, 4 CTA/SM", 4);
    run<2, 8>("DMMA x2 + DFMA x8, 4 CTA/SM", 4);
    return 0;
}
