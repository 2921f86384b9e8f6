// FP64 roofline denominator: a DFMA-only microbenchmark (independent chains, no memory
// traffic). MEASURED_PEAKS.json carries HBM and bf16 figures only, and the propagate path is
// bound by the FP64 pipe, so bench.py measures this peak in the same
// run, under the same clocks, and quotes fractions of it next to the nominal
// 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s.
#include <cuda_runtime.h>

#include "../../include/perturb_gpu.h"

namespace {

constexpr int kChains = 8;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b) {
    double x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = a + c * 1e-3 + threadIdx.x * 1e-6;
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = __fma_rn(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the loop
}

}  // namespace

extern "C" int perturb_gpu_measure_fp64_peak(int device, double *tflops) {
    if (tflops == nullptr) return PERTURB_GPU_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return PERTURB_GPU_ERR_NODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PERTURB_GPU_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8 * 4;
    const int threads = 256;
    double *d_out = nullptr;
    if (cudaMalloc(&d_out, sizeof(double) * blocks * threads) != cudaSuccess) {
        return PERTURB_GPU_ERR_ALLOC;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(d_out, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * kChains * static_cast<double>(kIters) * blocks * threads;
        if (rep >= 1 && ms > 0.f) best = best > flops / (ms * 1e-3) ? best : flops / (ms * 1e-3);
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (prev >= 0) cudaSetDevice(prev);
    if (err != cudaSuccess) return PERTURB_GPU_ERR_CUDA;
    *tflops = best * 1e-12;
    return PERTURB_GPU_OK;
}
