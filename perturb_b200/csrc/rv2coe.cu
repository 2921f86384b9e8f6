// N4: classical orbital elements from TEME states, batched -- the
// counterpart of perturb::ClassicalOrbitalElements(StateVector, GravModel)
// (src/perturb.cpp:119-131) and sgp4::rv2coe_SGP4 / newtonnu_SGP4 / angle_SGP4
// (src/sgp4.cpp:2863-3064, 2749-2803, 2659-2681), which the reference's verification
// printout applies to every propagated state (tests/test_perturb.cpp:385-397).
//
// Element-wise over (satellite, time): 48 B read + 88 B written per state, so the kernel is
// HBM bound (6.5 TB/s / 136 B = 4.8e10 states/s). Compiled with -fmad=false: apart from
// libdevice-vs-glibc acos/atan2/sin/cos the arithmetic is the reference's, which matters for
// the 1e-8 "small" thresholds that switch outputs to the 999999.1 "undefined" sentinel.
#include <cuda_runtime.h>

#include "../../include/perturb_gpu.h"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kSmall = 0.00000001;
constexpr double kUndefined = 999999.1;
constexpr double kInfinite = 999999.9;

__device__ inline double mag3(const double v[3]) {
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

__device__ inline double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

__device__ inline double sgn(double x) { return x < 0.0 ? -1.0 : 1.0; }

// angle_SGP4, src/sgp4.cpp:2659-2681
__device__ inline double angle_between(const double a[3], const double b[3]) {
    const double ma = mag3(a), mb = mag3(b);
    if (ma * mb > kSmall * kSmall) {
        double temp = dot3(a, b) / (ma * mb);
        if (fabs(temp) > 1.0) temp = sgn(temp) * 1.0;
        return acos(temp);
    }
    return kUndefined;
}

// newtonnu_SGP4, src/sgp4.cpp:2749-2803
__device__ inline void newtonnu(double ecc, double nu, double &e0, double &m) {
    e0 = 999999.9;
    m = 999999.9;
    if (fabs(ecc) < kSmall) {
        m = nu;
        e0 = nu;
    } else if (ecc < 1.0 - kSmall) {
        const double sine = (sqrt(1.0 - ecc * ecc) * sin(nu)) / (1.0 + ecc * cos(nu));
        const double cose = (ecc + cos(nu)) / (1.0 + ecc * cos(nu));
        e0 = atan2(sine, cose);
        m = e0 - ecc * sin(e0);
    } else if (ecc > 1.0 + kSmall) {
        if ((ecc > 1.0) && (fabs(nu) + 0.00001 < kPi - acos(1.0 / ecc))) {
            const double sine = (sqrt(ecc * ecc - 1.0) * sin(nu)) / (1.0 + ecc * cos(nu));
            e0 = log(sine + sqrt(sine * sine + 1.0));  // asinh_SGP4, :2705-2711
            m = ecc * sinh(e0) - e0;
        }
    } else if (fabs(nu) < 168.0 * kPi / 180.0) {
        e0 = tan(nu * 0.5);
        m = e0 + (e0 * e0 * e0) / 3.0;
    }
    if (ecc < 1.0) {
        m = fmod(m, 2.0 * kPi);
        if (m < 0.0) m = m + 2.0 * kPi;
        e0 = fmod(e0, 2.0 * kPi);
    }
}

// rv2coe_SGP4, src/sgp4.cpp:2863-3064. out[0..10] = p, a, ecc, incl, omega, argp, nu, m,
// arglat, truelon, lonper.
__device__ inline void rv2coe(double r[3], double v[3], double mus, double out[11]) {
    double p, a, ecc, incl, omega, argp, nu, m, arglat, truelon, lonper;
    const double halfpi = 0.5 * kPi;
    const double magr = mag3(r), magv = mag3(v);
    double hbar[3] = {r[1] * v[2] - r[2] * v[1], r[2] * v[0] - r[0] * v[2],
                      r[0] * v[1] - r[1] * v[0]};
    const double magh = mag3(hbar);
    if (magh > kSmall) {
        double nbar[3] = {-hbar[1], hbar[0], 0.0};
        const double magn = mag3(nbar);
        const double c1 = magv * magv - mus / magr;
        const double rdotv = dot3(r, v);
        double ebar[3];
        for (int i = 0; i <= 2; i++) ebar[i] = (c1 * r[i] - rdotv * v[i]) / mus;
        ecc = mag3(ebar);

        const double sme = (magv * magv * 0.5) - (mus / magr);
        if (fabs(sme) > kSmall) {
            a = -mus / (2.0 * sme);
        } else {
            a = kInfinite;
        }
        p = magh * magh / mus;

        const double hk = hbar[2] / magh;
        incl = acos(hk);

        // 1 = elliptical/parabolic/hyperbolic inclined, 2 = circular equatorial,
        // 3 = circular inclined, 4 = elliptical equatorial
        int typeorbit = 1;
        const bool equatorial = (incl < kSmall) | (fabs(incl - kPi) < kSmall);
        if (ecc < kSmall) {
            typeorbit = equatorial ? 2 : 3;
        } else if (equatorial) {
            typeorbit = 4;
        }

        if (magn > kSmall) {
            double temp = nbar[0] / magn;
            if (fabs(temp) > 1.0) temp = sgn(temp);
            omega = acos(temp);
            if (nbar[1] < 0.0) omega = kTwoPi - omega;
        } else {
            omega = kUndefined;
        }

        if (typeorbit == 1) {
            argp = angle_between(nbar, ebar);
            if (ebar[2] < 0.0) argp = kTwoPi - argp;
        } else {
            argp = kUndefined;
        }

        if ((typeorbit == 1) || (typeorbit == 4)) {
            nu = angle_between(ebar, r);
            if (rdotv < 0.0) nu = kTwoPi - nu;
        } else {
            nu = kUndefined;
        }

        m = 0.0;
        if (typeorbit == 3) {
            arglat = angle_between(nbar, r);
            if (r[2] < 0.0) arglat = kTwoPi - arglat;
            m = arglat;
        } else {
            arglat = kUndefined;
        }

        if ((ecc > kSmall) && (typeorbit == 4)) {
            double temp = ebar[0] / ecc;
            if (fabs(temp) > 1.0) temp = sgn(temp);
            lonper = acos(temp);
            if (ebar[1] < 0.0) lonper = kTwoPi - lonper;
            if (incl > halfpi) lonper = kTwoPi - lonper;
        } else {
            lonper = kUndefined;
        }

        if ((magr > kSmall) && (typeorbit == 2)) {
            double temp = r[0] / magr;
            if (fabs(temp) > 1.0) temp = sgn(temp);
            truelon = acos(temp);
            if (r[1] < 0.0) truelon = kTwoPi - truelon;
            if (incl > halfpi) truelon = kTwoPi - truelon;
            m = truelon;
        } else {
            truelon = kUndefined;
        }

        if ((typeorbit == 1) || (typeorbit == 4)) {
            double e;
            newtonnu(ecc, nu, e, m);
        }
    } else {
        p = a = ecc = incl = omega = argp = nu = m = arglat = truelon = lonper = kUndefined;
    }
    out[0] = p;
    out[1] = a;
    out[2] = ecc;
    out[3] = incl;
    out[4] = omega;
    out[5] = argp;
    out[6] = nu;
    out[7] = m;
    out[8] = arglat;
    out[9] = truelon;
    out[10] = lonper;
}

__global__ void __launch_bounds__(256)
rv2coe_kernel(const double *__restrict__ posvel, size_t n, size_t m, size_t row_stride,
              size_t plane_stride, double mus, double *__restrict__ coe, size_t coe_plane_stride) {
    const size_t kblocks = (m + blockDim.x - 1) / blockDim.x;
    const size_t i = blockIdx.x / kblocks;
    const size_t k = (blockIdx.x % kblocks) * blockDim.x + threadIdx.x;
    if (k >= m || i >= n) return;
    const size_t at = i * row_stride + k;  // lanes = consecutive times: coalesced planes
    double r[3], v[3], out[11];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = posvel[c * plane_stride + at];
        v[c] = posvel[(c + 3) * plane_stride + at];
    }
    const bool have = !(isnan(r[0]) || isnan(v[0]));  // codes 1-4 carry no state
    if (have) {
        rv2coe(r, v, mus, out);
    } else {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int c = 0; c < 11; ++c) out[c] = nan;
    }
#pragma unroll
    for (int c = 0; c < 11; ++c) coe[c * coe_plane_stride + at] = out[c];
}

bool on_device(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

}  // namespace

extern "C" int perturb_gpu_rv2coe(const double *posvel, size_t n, size_t m, size_t row_stride,
                                  size_t plane_stride, int grav_model, double *coe,
                                  size_t coe_plane_stride, void *stream) {
    if (posvel == nullptr || coe == nullptr || row_stride < m || plane_stride < n * row_stride ||
        coe_plane_stride < n * row_stride || n * ((m + 255) / 256) > 2147483647ULL) {
        return PERTURB_GPU_ERR_ARG;
    }
    double mus;  // getgravconst, src/sgp4.cpp:2119,2130,2141
    switch (grav_model) {
        case PERTURB_GPU_WGS72_OLD: mus = 398600.79964; break;
        case PERTURB_GPU_WGS72: mus = 398600.8; break;
        case PERTURB_GPU_WGS84: mus = 398600.5; break;
        default: return PERTURB_GPU_ERR_ARG;
    }
    if (n == 0 || m == 0) return PERTURB_GPU_OK;
    if (!on_device(posvel) || !on_device(coe)) return PERTURB_GPU_ERR_ARG;  // device buffers only
    const unsigned grid = static_cast<unsigned>(n * ((m + 255) / 256));
    rv2coe_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        posvel, n, m, row_stride, plane_stride, mus, coe, coe_plane_stride);
    return cudaGetLastError() == cudaSuccess ? PERTURB_GPU_OK : PERTURB_GPU_ERR_CUDA;
}
