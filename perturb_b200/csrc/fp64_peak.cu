// FP64 roofline denominator: a DFMA-only microbenchmark (independent chains, no memory
// traffic). MEASURED_PEAKS.json carries HBM and bf16 figures only, and the propagate path is
// bound by the FP64 pipe, so bench.py measures this peak in the same
// run, under the same clocks, and quotes fractions of it next to the nominal
// 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s.
#include <cuda_runtime.h>
#include <time.h>

#include "../../include/perturb_gpu.h"
#include "../../include/perturb_benchtools.h"

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

// Raw D2H rate into pinned host memory: the ceiling of every host-buffer (end-to-end) figure.
extern "C" int perturb_gpu_measure_d2h(int device, size_t bytes, int reps, double *gbs_best,
                                       double *gbs_mean) {
    if (gbs_best == nullptr || bytes == 0 || reps < 1) return PERTURB_GPU_ERR_ARG;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return PERTURB_GPU_ERR_NODEV;
    void *d = nullptr, *h = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int status = PERTURB_GPU_OK;
    double best = 0.0, total_ms = 0.0;
    if (cudaMalloc(&d, bytes) != cudaSuccess || cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess) {
        status = PERTURB_GPU_ERR_ALLOC;
    } else if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess ||
               cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        status = PERTURB_GPU_ERR_CUDA;
    } else {
        cudaMemsetAsync(d, 1, bytes, stream);
        cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream);  // warm-up: first touch
        cudaStreamSynchronize(stream);
        for (int r = 0; r < reps; ++r) {
            cudaEventRecord(e0, stream);
            cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream);
            cudaEventRecord(e1, stream);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms > 0.f) {
                total_ms += ms;
                const double g = static_cast<double>(bytes) / (ms * 1e-3) * 1e-9;
                best = g > best ? g : best;
            }
        }
        if (cudaGetLastError() != cudaSuccess) status = PERTURB_GPU_ERR_CUDA;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (stream) cudaStreamDestroy(stream);
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    if (prev >= 0) cudaSetDevice(prev);
    if (status != PERTURB_GPU_OK) return status;
    *gbs_best = best;
    if (gbs_mean != nullptr) {
        *gbs_mean = total_ms > 0.0 ? static_cast<double>(bytes) * reps / (total_ms * 1e-3) * 1e-9 : 0.0;
    }
    return PERTURB_GPU_OK;
}

// The same from ONE process over several devices at once: every device copies `bytes` bytes
// `reps` times into its own pinned buffer (allocated with `host_flags`: 0 default,
// 1 cudaHostAllocPortable, 4 cudaHostAllocWriteCombined, or their sum), all copies in flight
// together; *gbs_total = all bytes / wall time between the first enqueue and the last
// completion. With the per-process figure of perturb_gpu_measure_d2h this separates what the
// box can land in host memory from how the buffers were allocated.
extern "C" int perturb_gpu_measure_d2h_multi(int n_devices, size_t bytes, int reps,
                                             unsigned host_flags, double *gbs_total) {
    if (gbs_total == nullptr || bytes == 0 || reps < 1 || n_devices < 1 || n_devices > 64) {
        return PERTURB_GPU_ERR_ARG;
    }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < n_devices) return PERTURB_GPU_ERR_NODEV;
    int prev = -1;
    cudaGetDevice(&prev);
    void *d[64] = {nullptr}, *h[64] = {nullptr};
    cudaStream_t st[64] = {nullptr};
    int status = PERTURB_GPU_OK;
    for (int g = 0; g < n_devices && status == PERTURB_GPU_OK; ++g) {
        if (cudaSetDevice(g) != cudaSuccess || cudaMalloc(&d[g], bytes) != cudaSuccess ||
            cudaHostAlloc(&h[g], bytes, host_flags) != cudaSuccess ||
            cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking) != cudaSuccess) {
            status = PERTURB_GPU_ERR_ALLOC;
            break;
        }
        cudaMemsetAsync(d[g], 1, bytes, st[g]);
        cudaMemcpyAsync(h[g], d[g], bytes, cudaMemcpyDeviceToHost, st[g]);  // warm-up / first touch
    }
    if (status == PERTURB_GPU_OK) {
        for (int g = 0; g < n_devices; ++g) {
            cudaSetDevice(g);
            cudaStreamSynchronize(st[g]);
        }
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int r = 0; r < reps; ++r) {
            for (int g = 0; g < n_devices; ++g) {
                cudaSetDevice(g);
                cudaMemcpyAsync(h[g], d[g], bytes, cudaMemcpyDeviceToHost, st[g]);
            }
        }
        for (int g = 0; g < n_devices; ++g) {
            cudaSetDevice(g);
            cudaStreamSynchronize(st[g]);
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        *gbs_total = static_cast<double>(bytes) * reps * n_devices / secs * 1e-9;
        if (cudaGetLastError() != cudaSuccess) status = PERTURB_GPU_ERR_CUDA;
    }
    for (int g = 0; g < n_devices; ++g) {
        cudaSetDevice(g);
        if (st[g]) cudaStreamDestroy(st[g]);
        if (h[g]) cudaFreeHost(h[g]);
        if (d[g]) cudaFree(d[g]);
    }
    cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
    return status;
}
