// FP64 latency / throughput microbenchmark for sm_100a: dependent DFMA chains, ILP chains per thread, W warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warpsPerSM) {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<ILP><<<148, warpsPerSM * 32>>>(out, 10, 1.0000001, 1e-9, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP><<<148, warpsPerSM * 32>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double instrPerWarp = (double)iters * 8 * ILP;
    printf("ILP %2d warps/SM %2d: %.2f cycles per DFMA per warp, SM-wide %.2f DFMA(warp)/cycle, %.1f TFLOP/s\n", ILP, warpsPerSM,
           c / instrPerWarp, instrPerWarp * warpsPerSM / c, 2.0 * instrPerWarp * 32 * warpsPerSM * 148 / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
