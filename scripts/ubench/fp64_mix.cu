// Micro-benchmark: does an FP64 instruction block the issue port for its 2 pipe cycles?
// Times per-thread loops of DFMA only, DFMA + k independent integer ops, and integer only.
#include <cstdio>
#include <cuda_runtime.h>

template <int K_INT, bool WITH_FP64>
__global__ void __launch_bounds__(256) mix(double *out, int *iout, double a, double b, int iters) {
    double x0 = a + threadIdx.x, x1 = a * 2, x2 = a * 3, x3 = a * 4;
    int i0 = threadIdx.x, i1 = 2, i2 = 3, i3 = 5;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (WITH_FP64) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            }
#pragma unroll
            for (int k = 0; k < K_INT; ++k) {
                i0 = (i0 ^ i1) + 0x9e37; i1 = (i1 ^ i2) + 0x79b9; i2 = (i2 ^ i3) + 0x7f4a; i3 = (i3 ^ i0) + 0x7c15;
            }
        }
    }
    if (x0 + x1 + x2 + x3 == 1.2345) out[threadIdx.x] = x0;
    if ((i0 ^ i1 ^ i2 ^ i3) == 0x1234567) iout[threadIdx.x] = i0;
}

template <int K_INT, bool WITH_FP64>
double run(const char *name) {
    double *d; int *di;
    cudaMalloc(&d, 4096); cudaMalloc(&di, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, blocks = 148 * 8;
    mix<K_INT, WITH_FP64><<<blocks, 256>>>(d, di, 0.999, 1e-9, iters);
    cudaEventRecord(e0);
    mix<K_INT, WITH_FP64><<<blocks, 256>>>(d, di, 0.999, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_dfma = WITH_FP64 ? 32.0 * iters : 0.0;       // per warp
    const double warp_int = 2.0 * 4 * 8.0 * K_INT * iters;          // LOP3 + IADD per statement
    printf("%-28s %8.3f ms   fp64 warp-instr %.0f  int warp-instr %.0f\n", name, ms, warp_dfma, warp_int);
    cudaFree(d); cudaFree(di);
    return ms;
}

int main() {
    const double f = run<0, true>("DFMA only");
    const double i1 = run<1, false>("INT x1 only");
    const double i2 = run<2, false>("INT x2 only");
    const double m1 = run<1, true>("DFMA + INT x1 (1:2 instr)");
    const double m2 = run<2, true>("DFMA + INT x2 (1:4 instr)");
    printf("if co-issue were free: max(f, i). measured/ideal: x1 %.2f  x2 %.2f ; additive model f+i: x1 %.2f x2 %.2f\n",
           m1 / (f > i1 ? f : i1), m2 / (f > i2 ? f : i2), m1 / (f + i1), m2 / (f + i2));
    return 0;
}
