// Test hook: evaluates the device math routines of sgp4_math.cuh element-wise so that
// tests/test_gpu_math.py can measure their error against correctly-rounded host values.
#include <cuda_runtime.h>

#include "../../include/perturb_gpu.h"
#include "../../include/perturb_benchtools.h"
#include "sgp4_math.cuh"

namespace {

__global__ void probe_kernel(int fn, const double *x, const double *y, double *out, size_t n) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const double a = x[i], b = y[i];
    double s = 0.0, c = 0.0, r = 0.0;
    switch (fn) {
        case 0: pb::sincos_cw(a, s, c); r = s; break;
        case 1: pb::sincos_cw(a, s, c); r = c; break;
        case 2: r = pb::sin_cw(a); break;
        case 3: r = pb::cos_cw(a); break;
        case 4: r = pb::div_fast(a, b); break;
        case 5: r = pb::rcp_fast(a); break;
        case 6: r = pb::sqrt_fast(a); break;
        case 7: r = pb::rsqrt_fast(a); break;
        case 8: r = pb::fmod_2pi(a); break;
        case 9: r = pb::wrap_pi(a); break;
        case 10: r = pb::wrap_pi2(a); break;
        case 11: pb::sincos_small(a, s, c); r = s; break;
        case 12: pb::sincos_small(a, s, c); r = c; break;
        case 13: r = pb::pow_2o3(a); break;
        case 14: r = pb::div_coarse(a, b); break;
        case 15: pb::sincos_small<9>(a, s, c); r = s; break;
        case 16: pb::sincos_small<9>(a, s, c); r = c; break;
        default: r = 0.0; break;
    }
    out[i] = r;
}

}  // namespace

extern "C" int perturb_gpu_math_probe(int fn, const double *x, const double *y, double *out,
                                      size_t n) {
    if (x == nullptr || y == nullptr || out == nullptr) return PERTURB_GPU_ERR_ARG;
    if (n == 0) return PERTURB_GPU_OK;
    double *d = nullptr;
    if (cudaMalloc(&d, 3 * n * sizeof(double)) != cudaSuccess) return PERTURB_GPU_ERR_ALLOC;
    cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, y, n * sizeof(double), cudaMemcpyHostToDevice);
    probe_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(fn, d, d + n, d + 2 * n, n);
    const cudaError_t e1 = cudaMemcpy(out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    const cudaError_t e2 = cudaGetLastError();
    cudaFree(d);
    return (e1 == cudaSuccess && e2 == cudaSuccess) ? PERTURB_GPU_OK : PERTURB_GPU_ERR_CUDA;
}
