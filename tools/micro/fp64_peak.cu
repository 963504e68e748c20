// Measured FP64 peak of the device ("measure with a DFMA-chain microbenchmark on the box and record it"):
// independent DFMA chains, 8 per thread, 16 and 32 warps per SM, best of 5 launches.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak > profiles/r02_fp64_peak.json
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, iters = 4000;
    double* out;
    cudaMalloc(&out, (size_t)sms * 1024 * 8);
    double best = 0.0;
    int bestWarps = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int warps : {16, 32}) {
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            k<8><<<sms, warps * 32>>>(out, rep ? iters : 10, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (!rep) continue;
            const double tflops = 2.0 * (double)iters * 8 * 8 * 32 * warps * sms / (ms * 1e-3) / 1e12;
            if (tflops > best) best = tflops, bestWarps = warps;
        }
    }
    int clockKHz = 0;
    cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0);
    printf("{\"fp64_tflops\": %.2f, \"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"warps_per_sm\": %d, "
           "\"dfma_per_clk_per_sm\": %.1f, \"how\": \"tools/micro/fp64_peak.cu: 8 independent DFMA chains per thread, best of 5 launches\"}\n",
           best, prop.name, sms, clockKHz / 1e3, bestWarps, best * 1e12 / 2.0 / sms / (clockKHz * 1e3));
    return 0;
}
