// Microbenchmark: DFMA throughput on sm_100a as a function of warps per scheduler and independent chains per thread
// (how much ILP the Fisher kernel's two warps per scheduler need to keep the FP64 pipe busy).
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void k(double* out, int iters, double a, double b) {
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += x[i];
    if (s == 1234.5) out[0] = s;
}
template <int N> void run(int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    k<N><<<sms, warps_per_sm * 32>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); k<N><<<sms, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)sms * warps_per_sm * iters * 16.0 * N;            // warp instructions
    const double per_clk_smsp = inst / (ms * 1e-3 * 1.965e9 * sms * 4);
    printf("warps/SM %2d chains %d: %.3f DFMA warp-instr / clk / scheduler (peak 0.5) -> %.1f%%; cycles per dependent DFMA per warp %.1f\n", warps_per_sm, N,
           per_clk_smsp, 200 * per_clk_smsp, (warps_per_sm / 4.0) * N / per_clk_smsp / N);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 12, 16}) {
        run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w);
    }
    return 0;
}
