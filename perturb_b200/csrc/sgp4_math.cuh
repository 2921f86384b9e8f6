// Device FP64 math for the SGP4/SDP4 kernels (sm_100a).
//
// The reference calls glibc's sin/cos/fmod/pow/atan2 (L0 of the layer map). On B200 the
// path is FP64-pipe bound, so the kernels use range-bounded fast paths:
//   * sincos: 3-constant Cody-Waite reduction + minimax kernels on [-pi/4, pi/4], with a
//     (warp-uniformly not taken) fall-back to libdevice for |x| >= 1e5;
//   * fmod(x, 2pi): one FMA-exact reduction with the C `fmod` sign convention (libdevice's
//     fmod is a bit-serial integer loop: 0 FP64 but ~144 integer instructions);
//   * pow(x, 2/3) -> cbrt(x)^2, pow(x, 1.5) -> x*sqrt(x).
// None of these is bit-identical to glibc; the parity budget (1e-6 km, 1e-9 km/s) leaves
// about five decimal orders over their rounding error (tests/test_gpu_parity.py).
#ifndef PERTURB_B200_SGP4_MATH_CUH
#define PERTURB_B200_SGP4_MATH_CUH

#include <cuda_runtime.h>

namespace pb {

constexpr double kPi = 3.14159265358979323846;  // src/sgp4.cpp:69
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kInvTwoPi = 1.0 / kTwoPi;
constexpr double kRintMagic = 6755399441055744.0;  // 1.5 * 2^52

// x mod 2pi with the sign of x (C fmod semantics), exact: q = rint(x/2pi) may be off by one
// from trunc(x/2pi), but x - q*2pi is a multiple of 2^-50 below 8 in magnitude, so the FMA
// and the +-2pi fix-ups are all exact and the result is the unique representative fmod returns.
__device__ __forceinline__ double fmod_2pi(double x) {
    const double q = __dadd_rn(__fma_rn(x, kInvTwoPi, kRintMagic), -kRintMagic);
    double r = __fma_rn(-q, kTwoPi, x);
    if (x >= 0.0) {
        if (r < 0.0) r = __dadd_rn(r, kTwoPi);
    } else {
        if (r > 0.0) r = __dadd_rn(r, -kTwoPi);
    }
    return r;
}

// Same reduction without the sign fix-up: result in [-pi, pi], congruent to x mod 2pi.
// For consumers that only feed the angle to sin/cos or to sums that are reduced again.
__device__ __forceinline__ double wrap_pi(double x) {
    const double q = __dadd_rn(__fma_rn(x, kInvTwoPi, kRintMagic), -kRintMagic);
    return __fma_rn(-q, kTwoPi, x);
}

namespace detail {
// minimax kernels on [-pi/4, pi/4] (the classic fdlibm coefficient sets)
__device__ __forceinline__ double sin_kernel(double r, double z) {
    double p = 1.58969099521155010221e-10;
    p = __fma_rn(p, z, -2.50507602534068634195e-08);
    p = __fma_rn(p, z, 2.75573137070700676789e-06);
    p = __fma_rn(p, z, -1.98412698298579493134e-04);
    p = __fma_rn(p, z, 8.33333333332248946124e-03);
    p = __fma_rn(p, z, -1.66666666666666324348e-01);
    return __fma_rn(__dmul_rn(r, z), p, r);
}
__device__ __forceinline__ double cos_kernel(double z) {
    double p = -1.13596475577881948265e-11;
    p = __fma_rn(p, z, 2.08757232129817482790e-09);
    p = __fma_rn(p, z, -2.75573143513906633035e-07);
    p = __fma_rn(p, z, 2.48015872894767294178e-05);
    p = __fma_rn(p, z, -1.38888888888741095749e-03);
    p = __fma_rn(p, z, 4.16666666666666019037e-02);
    return __fma_rn(__dmul_rn(z, z), p, __fma_rn(-0.5, z, 1.0));
}
}  // namespace detail

// sin and cos of x together. Fast path valid for |x| < 1e5 (k < 2^17, so the three-term
// pi/2 expansion leaves < 1e-27 absolute error in the reduced argument).
__device__ __forceinline__ void sincos_fast(double x, double &s, double &c) {
    if (!(fabs(x) < 1.0e5)) {  // also catches NaN
        sincos(x, &s, &c);
        return;
    }
    const double kd = __fma_rn(x, 6.36619772367581382433e-01, kRintMagic);
    const int q = __double2loint(kd);
    const double k = __dadd_rn(kd, -kRintMagic);
    double r = __fma_rn(-k, 1.5707963267948966e+00, x);
    r = __fma_rn(-k, 6.1232339957367574e-17, r);
    r = __fma_rn(-k, 8.4784276603688985e-32, r);
    const double z = __dmul_rn(r, r);
    const double sr = detail::sin_kernel(r, z);
    const double cr = detail::cos_kernel(z);
    double ss = (q & 1) ? cr : sr;
    double cc = (q & 1) ? sr : cr;
    if (q & 2) ss = -ss;
    if ((q + 1) & 2) cc = -cc;
    s = ss;
    c = cc;
}

__device__ __forceinline__ double sin_fast(double x) {
    double s, c;
    sincos_fast(x, s, c);
    return s;
}

__device__ __forceinline__ double cos_fast(double x) {
    double s, c;
    sincos_fast(x, s, c);
    return c;
}

// pow(x, 2/3) for x > 0 (the reference passes x2o3 = 2.0/3.0, src/sgp4.cpp:1795, 1867)
__device__ __forceinline__ double pow_2o3(double x) {
    const double c = cbrt(x);
    return c * c;
}

}  // namespace pb

#endif
