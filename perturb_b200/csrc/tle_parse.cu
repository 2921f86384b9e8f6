// TLE text front end: host entry point (perturb_gpu_parse_tles) and the device kernel behind
// perturb_gpu_init_from_text (sscanf costs 2.5-3 us per TLE on the
// host, i.e. seconds for a 1M-satellite catalogue, once propagation itself takes milliseconds).
#include <cuda_runtime.h>

#include <cmath>

#include "batch_internal.h"
#include "tle_parse_core.h"

namespace {

pb::tle::Pow10Table host_pow_table() {
    pb::tle::Pow10Table t;
    for (int i = 0; i < pb::tle::kPowCount; ++i) {
        // what both reference readers call (src/tle.cpp:136-137, src/sgp4.cpp:2337-2338)
        t.p[i] = std::pow(10.0, static_cast<double>(i - pb::tle::kPowBias));
    }
    return t;
}

__global__ void __launch_bounds__(128)
tle_parse_kernel(const char *__restrict__ lines1, const char *__restrict__ lines2, size_t stride,
                 int n, int flavor, const pb::tle::Pow10Table pw, perturb_gpu_tle *__restrict__ out,
                 int32_t *__restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // One thread per TLE; a warp reads 32 x 70 B of each line block, which is contiguous.
    char l1[72], l2[72];
    const char *g1 = lines1 + static_cast<size_t>(i) * stride;
    const char *g2 = lines2 + static_cast<size_t>(i) * stride;
#pragma unroll 1
    for (int c = 0; c < pb::tle::kLineLen; ++c) {
        l1[c] = g1[c];
        l2[c] = g2[c];
    }
    perturb_gpu_tle t;
    const int st = pb::tle::parse_one(l1, l2, flavor, pw, t);
    pb::tle::finish_entry(st, flavor, t);
    out[i] = t;
    if (status != nullptr) status[i] = st;
}

}  // namespace

namespace pb {

void launch_tle_parse(const char *d_lines1, const char *d_lines2, size_t stride, int n, int flavor,
                      void *d_tles, int32_t *d_status, cudaStream_t stream) {
    if (n <= 0) return;
    static const tle::Pow10Table pw = host_pow_table();
    tle_parse_kernel<<<(n + 127) / 128, 128, 0, stream>>>(
        d_lines1, d_lines2, stride, n, flavor, pw, static_cast<perturb_gpu_tle *>(d_tles), d_status);
}

}  // namespace pb

extern "C" int perturb_gpu_parse_tles(const char *lines1, const char *lines2,
                                      size_t line_stride, size_t n, int flavor,
                                      perturb_gpu_tle *out, int32_t *status) {
    if ((n > 0 && (lines1 == nullptr || lines2 == nullptr || out == nullptr)) ||
        line_stride < static_cast<size_t>(pb::tle::kLineLen) || (flavor != 0 && flavor != 1)) {
        return PERTURB_GPU_ERR_ARG;
    }
    static const pb::tle::Pow10Table pw = host_pow_table();
    for (size_t i = 0; i < n; ++i) {
        perturb_gpu_tle t;
        const int st = pb::tle::parse_one(lines1 + i * line_stride, lines2 + i * line_stride,
                                          flavor, pw, t);
        pb::tle::finish_entry(st, flavor, t);
        out[i] = t;
        if (status != nullptr) status[i] = st;
    }
    return PERTURB_GPU_OK;
}

extern "C" int perturb_gpu_parse_tle(const char *line_1, const char *line_2,
                                     perturb_gpu_tle *numbers, perturb_gpu_tle_ident *ident) {
    if (line_1 == nullptr || line_2 == nullptr || numbers == nullptr) return PERTURB_GPU_ERR_ARG;
    if (pb::tle::has_nul(line_1) || pb::tle::has_nul(line_2)) return 2;  // INVALID_FORMAT
    static const pb::tle::Pow10Table pw = host_pow_table();
    perturb_gpu_tle t;
    perturb_gpu_tle_ident id;
    const int st = pb::tle::parse_perturb(line_1, line_2, pw, t, &id);
    if (st == 0 || st >= 3) {
        *numbers = t;
        if (ident != nullptr) *ident = id;
    }
    return st;
}
