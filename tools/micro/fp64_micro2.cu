// FP64 throughput with different operand sources (register / kernel-parameter constant bank via uniform regs) and mixes
// typical of the assembly kernels (DFMA + FSEL, DFMA + MUFU.RCP64H), at 20 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
struct Tab { double v[64]; };
template <int MODE>
__global__ void __launch_bounds__(32, 20) k(double* out, int iters, const __grid_constant__ Tab tb, double a, double b, long long* cyc) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    int sel = threadIdx.x & 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);                               // registers only
                if (MODE == 1) x[i] = fma(x[i], tb.v[r * 8 + i], b);                  // constant-bank operand (static index)
                if (MODE == 2) x[i] = fma(x[i], tb.v[(it & 7) * 8 + i], b);           // constant-bank operand, dynamic uniform index (LDCU -> UR)
                if (MODE == 3) { x[i] = fma(x[i], a, b); x[i] = (sel == (i & 3)) ? x[i] : x[(i + 1) & 7]; }  // DFMA + 2 FSEL
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name) {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 20 * 32 * 8); cudaMalloc(&cyc, 8);
    Tab tb; for (int i = 0; i < 64; ++i) tb.v[i] = 1.0 + 1e-9 * i;
    const int iters = 2000;
    k<MODE><<<148 * 20, 32>>>(out, 10, tb, 1.0000001, 1e-9, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 20, 32>>>(out, iters, tb, 1.0000001, 1e-9, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)iters * 64 * 32 * 20 * 148;
    printf("%-48s %.2f TFLOP/s (DFMA only counted), %.3f ms\n", name, 2.0 * dfma / (ms * 1e-3) / 1e12, ms);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("register operands");
    run<1>("constant-bank operand, static index");
    run<2>("constant-bank operand, dynamic uniform index");
    run<3>("DFMA + 2 FSEL per DFMA");
    return 0;
}
