// Device FP64 math for the SGP4/SDP4 kernels (sm_100a).
//
// The reference calls glibc's sin/cos/fmod/pow/atan2 (L0 of the layer map). On B200 the
// path is bound by the FP64 pipe AND by issue slots: an FP64 instruction holds its
// scheduler's issue port for both pipe cycles (scripts/ubench/fp64_mix.cu), so every
// non-FP64 instruction costs time as well. The first version spent ~1475 warp instructions
// per evaluation of which ~560 were FP64 (ncu): constant materialisation (two UMOV per
// double literal), libdevice slow-path guards around div/sqrt/sincos, quadrant logic. Hence:
//   * polynomial and reduction constants, and every literal of the path that is not exact in
//     32 bits, live in one __constant__ table, fetched two at a time into uniform registers
//     (LDCU.128) instead of two moves each;
//   * sincos: two-constant Cody-Waite reduction by pi/2 (FMA-exact first step, good for any
//     angle the path produces) + minimax kernels on [-pi/4, pi/4] in Horner form (every
//     coefficient but the leading one is a uniform-register operand: at 28 warps per SM the
//     kernel is issue bound, not latency bound, and Estrin's extra multiplies and loads of
//     coefficient pairs cost more than its shorter dependency chain saves), no slow path;
//     small angles use short Taylor sums (sincos_small, sincos_taylor) and rotations;
//   * fmod(x, 2pi): one FMA-exact reduction with the C `fmod` sign convention (libdevice's
//     fmod is a bit-serial integer loop); wrap_pi where only the angle mod 2pi matters;
//   * division / sqrt: MUFU seed + one third-order (reciprocal) or two Newton (rsqrt) steps,
//     no denormal/overflow fix-up call, selects instead of branches;
//   * pow(x, 2/3) -> cbrt(x)^2, pow(x, 1.5) -> x*sqrt(x).
// None of these is bit-identical to glibc; the parity budget (1e-6 km, 1e-9 km/s) leaves
// about two decimal orders over their rounding error (tests/test_gpu_parity.py); their own
// error bounds are tested in tests/test_gpu_math.py.
#ifndef PERTURB_B200_SGP4_MATH_CUH
#define PERTURB_B200_SGP4_MATH_CUH

#include <cuda_runtime.h>

namespace pb {

constexpr double kPi = 3.14159265358979323846;  // src/sgp4.cpp:69
constexpr double kTwoPi = 2.0 * kPi;
// 1.5 * 2^52: adding it rounds to the nearest integer. Its low word is zero, so as a literal it
// is an FP64 immediate operand (from the constant table it cost a load into a vector register
// per use).
constexpr double kRintMagic = 6755399441055744.0;

// Constant-bank table (index names below), one array so that neighbours load pairwise.
enum {
    MC_INV_TWOPI = 0,
    MC_TWOPI,
    MC_TWOPI_LO,        // 2pi - (double)2pi
    MC_TWO_OVER_PI,
    MC_PIO2_HI,
    MC_PIO2_MID,
    MC_S1, MC_S2, MC_S3, MC_S4, MC_S5, MC_S6,  // sin kernel, fdlibm minimax set
    MC_C1, MC_C2, MC_C3, MC_C4, MC_C5, MC_C6,  // cos kernel
    MC_HALF,
    MC_ONE,
    MC_T3, MC_T5, MC_T7, MC_T9,          // Taylor sin: -1/6, 1/120, -1/5040, 1/362880
    MC_U2, MC_U4, MC_U6, MC_U8, MC_U10,  // Taylor cos: -1/2, 1/24, -1/720, 1/40320, -1/3628800
    // literals of sgp4()/dspace()/dpper() that are not exact in 32 bits: as immediates each
    // costs two moves per use, as constant-bank operands nothing
    MC_PI, MC_1EM6, MC_1EM3, MC_0P95, MC_1EM12, MC_RPTIM, MC_ZNS, MC_2ZES, MC_ZNL, MC_2ZEL,
    MC_0P2, MC_1P5EM12, MC_STEP2,
    // sin / cos of the resonance phase constants, src/sgp4.cpp:1036-1043
    MC_S_FASX2, MC_C_FASX2, MC_S_FASX4, MC_C_FASX4, MC_S_FASX6, MC_C_FASX6, MC_S_G22, MC_C_G22,
    MC_S_G32, MC_C_G32, MC_S_G44, MC_C_G44, MC_S_G52, MC_C_G52, MC_S_G54, MC_C_G54,
    // atan(t) = t + t^3 P(t^2) on |t| <= 0.25 (scripts/fit_atan.py 0.2501 8: 1 ulp)
    MC_A0, MC_A1, MC_A2, MC_A3, MC_A4, MC_A5, MC_A6, MC_A7,
    // (1 + x)^(-2/3) = 1 + x (B1 + x (B2 + ...)): binomial coefficients -2/3, 5/9, -40/81, ...
    MC_B1, MC_B2, MC_B3, MC_B4, MC_B5, MC_B6,
    MC_KEPLER_SMALL, MC_5O3,  // kepler_small: its e^2 bound, 5/3
    MC_COUNT
};

static __constant__ double kMc[MC_COUNT] = {
    0.15915494309189534561,
    6.28318530717958647692,
    2.4492935982947064e-16,
    6.36619772367581382433e-01,
    1.5707963267948966e+00,
    6.1232339957367574e-17,
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    0.5,
    1.0,
    -1.66666666666666666667e-01, 8.33333333333333333333e-03, -1.98412698412698412698e-04,
    2.75573192239858906526e-06,
    -0.5, 4.16666666666666666667e-02, -1.38888888888888888889e-03, 2.48015873015873015873e-05,
    -2.75573192239858906526e-07,
    3.14159265358979323846, 1.0e-6, 0.001, 0.95, 1.0e-12, 4.37526908801129966e-3, 1.19459e-5,
    2.0 * 0.01675, 1.5835218e-4, 2.0 * 0.05490, 0.2, 1.5e-12, 259200.0,
    0.13093206501640100313, 0.99139134268488593686,   // fasx2 = 0.13130908
    0.25444411220371228015, -0.96708747989251968989,  // fasx4 = 2.8843198
    0.36578942533149936234, 0.93069763957778009002,   // fasx6 = 0.37448087
    -0.49213943048915524715, 0.87051638752972935101,  // g22 = 5.7686396
    0.81481440616389247062, 0.57972190187001152474,   // g32 = 0.95240898
    0.97350577801807993413, -0.22866241528815546759,  // g44 = 1.8014998
    0.86783740128127729936, 0.49684831179884195924,   // g52 = 1.0508330
    -0.95489237761530004228, -0.29695209575316896683, // g54 = 4.4108898
    -3.33333333333333037274e-01, 1.99999999999382532812e-01, -1.42857142649373713983e-01,
    1.11111084421623651508e-01, -9.09074023080557735987e-02, 7.68647824954676456288e-02,
    -6.55414157038025635416e-02, 4.72544858625266947505e-02,
    -6.66666666666666629659e-01, 5.55555555555555580227e-01, -4.93827160493827133081e-01,
    4.52674897119341557161e-01, -4.22496570644718794085e-01, 3.99024538942234441308e-01,
    4.0e-4, 1.66666666666666674068e+00,
};

// ---- warp votes ---------------------------------------------------------------------------
// Votes of the FULL warp as volatile inline PTX. Kernel 2 keeps its control flow warp-uniform
// with them (Kepler path choice, Newton loop exit, "every lane has a state to store"). The
// plain intrinsics are pure functions to the optimiser: in one instantiation (mean-element
// output, satellites split over the CTA's warps) it sank `__all_sync` below an `if (time index
// < m)` whose body was its only user, so the lanes past the end of the grid never reached the
// vote and the warp dead-locked. A volatile asm statement executes exactly where it is written.
__device__ __forceinline__ unsigned warp_ballot(bool p) {
    unsigned r;
    asm volatile(
        "{ .reg .pred q; setp.ne.u32 q, %1, 0; vote.sync.ballot.b32 %0, q, 0xffffffff; }"
        : "=r"(r)
        : "r"(static_cast<unsigned>(p)));
    return r;
}
__device__ __forceinline__ bool warp_all(bool p) { return warp_ballot(p) == 0xffffffffu; }
// Full-warp integer reductions (redux.sync), volatile for the same reason: their results feed
// stores that only lane 0 performs.
__device__ __forceinline__ unsigned warp_min_u32(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ unsigned warp_max_u32(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int warp_add_s32(int v) {
    int r;
    asm volatile("redux.sync.add.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}

// ---- reciprocal, division, square root ---------------------------------------------------
// MUFU.RCP64H / MUFU.RSQ64H seeds (about 20 good bits) refined in FP64. Inputs on
// this path are normal, finite doubles; zero / infinity / NaN propagate as NaN or infinity
// garbage exactly where the reference produces garbage too.
__device__ __forceinline__ double rcp_fast(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    // one third-order step: r (1 + e + e^2), e = 1 - b r  (20 -> 60 good bits)
    const double e = __fma_rn(-b, r, kMc[MC_ONE]);
    return __fma_rn(r, __fma_rn(e, e, e), r);
}

// a / b to ~2^-40 relative (seed + ONE Newton step, no residual correction). Only for
// quantities that are themselves Newton corrections, e.g. Kepler's equation: the iteration
// converges to the same root and its iterates move by < 1e-12 of each correction.
__device__ __forceinline__ double div_coarse(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = __fma_rn(r, __fma_rn(-b, r, kMc[MC_ONE]), r);
    return __dmul_rn(a, r);
}

__device__ __forceinline__ double div_fast(double a, double b) {
    const double r = rcp_fast(b);
    const double q = __dmul_rn(a, r);
    return __fma_rn(__fma_rn(-b, q, a), r, q);  // one correction step: <= 1 ulp
}

__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // y <- y * (1.5 - 0.5 x y^2), twice
    double h = __dmul_rn(x, kMc[MC_HALF]);
    double e = __fma_rn(-__dmul_rn(h, y), y, kMc[MC_HALF]);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-__dmul_rn(h, y), y, kMc[MC_HALF]);
    y = __fma_rn(y, e, y);
    return y;
}

__device__ __forceinline__ double sqrt_fast(double x) {
    const double y = rsqrt_fast(x);
    const double s = __dmul_rn(x, y);
    // s <- s + 0.5 y (x - s^2); x == 0 gives 0 * inf = NaN, patched by a select (no branch:
    // branches fence the uniform datapath)
    const double r = __fma_rn(__dmul_rn(y, kMc[MC_HALF]), __fma_rn(-s, s, x), s);
    return (x == 0.0) ? 0.0 : r;
}

// ---- angle reduction -----------------------------------------------------------------------
// x mod 2pi with the sign of x (C fmod semantics), exact: q = rint(x/2pi) may be off by one
// from trunc(x/2pi), but x - q*2pi is a multiple of 2^-50 below 8 in magnitude, so the FMA
// and the +-2pi fix-ups are all exact and the result is the unique representative fmod returns.
__device__ __forceinline__ double fmod_2pi(double x) {
    const double q = __dadd_rn(__fma_rn(x, kMc[MC_INV_TWOPI], kRintMagic),
                               -kRintMagic);
    const double r = __fma_rn(-q, kMc[MC_TWOPI], x);
    // one add of +2pi, -2pi or 0 chosen by selects: as nested branches this was two
    // reconvergence regions per call on the warp-uniform path
    const bool up = (x >= 0.0) && (r < 0.0);
    const bool down = !(x >= 0.0) && (r > 0.0);
    const double fix = up ? kMc[MC_TWOPI] : (down ? -kMc[MC_TWOPI] : 0.0);
    return __dadd_rn(r, fix);
}

// Same reduction modulo the DOUBLE 2pi without the sign fix-up: result in [-pi, pi]. For
// angles the reference reduces with fmod(x, twopi) and then only feeds to sin/cos or to sums
// that are reduced again (every near-Earth use).
__device__ __forceinline__ double wrap_pi(double x) {
    const double q = __dadd_rn(__fma_rn(x, kMc[MC_INV_TWOPI], kRintMagic),
                               -kRintMagic);
    return __fma_rn(-q, kMc[MC_TWOPI], x);
}

// Reduction modulo the TRUE 2pi (two-term constant): for angles the reference hands
// unreduced to sin/cos (xmdf and mm grow like n*t, ~700 rad after a week).
__device__ __forceinline__ double wrap_pi2(double x) {
    const double q = __dadd_rn(__fma_rn(x, kMc[MC_INV_TWOPI], kRintMagic),
                               -kRintMagic);
    const double r = __fma_rn(-q, kMc[MC_TWOPI], x);
    return __fma_rn(-q, kMc[MC_TWOPI_LO], r);
}

// ---- sin / cos ---------------------------------------------------------------------------
namespace detail {
// Quadrant k = rint(x * 2/pi) and reduced argument r = x - k*pi/2 with a two-term pi/2.
// x - k*PIO2_HI is a multiple of 2^-52 below 1 in magnitude, hence exact in one FMA for any k;
// the second term's rounding is half an ulp of r, and the 8.5e-32 k left out of pi/2 is
// below 1e-22 for |x| < 2^31: no prior reduction modulo 2 pi is needed for any angle of the
// path (mean anomalies reach 4e5 rad ten years from epoch). The quadrant comes from the low
// mantissa bits of the magic-number sum, correct for |k| < 2^51.
__device__ __forceinline__ double reduce_pio2(double x, int &q) {
    const double kd = __fma_rn(x, kMc[MC_TWO_OVER_PI], kRintMagic);
    q = __double2loint(kd);
    const double k = __dadd_rn(kd, -kRintMagic);
    double r = __fma_rn(-k, kMc[MC_PIO2_HI], x);
    r = __fma_rn(-k, kMc[MC_PIO2_MID], r);
    return r;
}
// Degree-5 polynomials in z, Horner form (fdlibm's minimax coefficients).
__device__ __forceinline__ double sin_poly(double z) {
    double p = kMc[MC_S6];
    p = __fma_rn(p, z, kMc[MC_S5]);
    p = __fma_rn(p, z, kMc[MC_S4]);
    p = __fma_rn(p, z, kMc[MC_S3]);
    p = __fma_rn(p, z, kMc[MC_S2]);
    return __fma_rn(p, z, kMc[MC_S1]);
}
__device__ __forceinline__ double cos_poly(double z) {
    double p = kMc[MC_C6];
    p = __fma_rn(p, z, kMc[MC_C5]);
    p = __fma_rn(p, z, kMc[MC_C4]);
    p = __fma_rn(p, z, kMc[MC_C3]);
    p = __fma_rn(p, z, kMc[MC_C2]);
    return __fma_rn(p, z, kMc[MC_C1]);
}
// v with its sign flipped when bit 1 of q is set (quadrants 2 and 3): one shift and one
// AND-XOR on the high word, integer pipe, not FP64.
__device__ __forceinline__ double flip_sign_q(double v, int q) {
    const int hi = __double2hiint(v) ^ ((q << 30) & 0x80000000);
    return __hiloint2double(hi, __double2loint(v));
}
}  // namespace detail

// sin and cos of any angle the path produces (|x| < 2^31); absolute error < 2.3e-16.
__device__ __forceinline__ void sincos_cw(double x, double &s, double &c) {
    int q;
    const double r = detail::reduce_pio2(x, q);
    const double z = __dmul_rn(r, r);
    const double sr = __fma_rn(__dmul_rn(r, z), detail::sin_poly(z), r);
    const double cr = __fma_rn(__fma_rn(detail::cos_poly(z), z, -kMc[MC_HALF]), z, kMc[MC_ONE]);
    const bool odd = (q & 1) != 0;
    const double ss = odd ? cr : sr;
    const double cc = odd ? sr : cr;
    s = detail::flip_sign_q(ss, q);
    c = detail::flip_sign_q(cc, q + 1);
}

// One of sin/cos only: which kernel is needed depends on quadrant parity, and both kernels
// have the same shape, so the coefficients are selected and ONE polynomial is run. Fewer FP64
// operations than sincos_cw but ~40 selects / constant loads: on B200 that is slower in the
// hot path (see the header), so the evaluation kernels use sincos_cw; kept for the probe.
template <bool WANT_COS>
__device__ __forceinline__ double sin_or_cos_cw(double x) {
    int q;
    const double r = detail::reduce_pio2(x, q);
    if (WANT_COS) q += 1;  // cos(x) = sin(x + pi/2)
    const bool use_cos = (q & 1) != 0;
    const double z = __dmul_rn(r, r);
    const double zz = __dmul_rn(z, z);
    const double a = __fma_rn(use_cos ? kMc[MC_C2] : kMc[MC_S2], z, use_cos ? kMc[MC_C1] : kMc[MC_S1]);
    const double b = __fma_rn(use_cos ? kMc[MC_C4] : kMc[MC_S4], z, use_cos ? kMc[MC_C3] : kMc[MC_S3]);
    const double c = __fma_rn(use_cos ? kMc[MC_C6] : kMc[MC_S6], z, use_cos ? kMc[MC_C5] : kMc[MC_S5]);
    const double p = __fma_rn(__fma_rn(c, zz, b), zz, a);
    // sin: r + (r z) p        cos: (1 - z/2) + (z z) p
    const double lead = use_cos ? __fma_rn(-kMc[MC_HALF], z, kMc[MC_ONE]) : r;
    const double scale = use_cos ? zz : __dmul_rn(r, z);
    const double v = __fma_rn(scale, p, lead);
    return detail::flip_sign_q(v, q);
}

__device__ __forceinline__ double sin_cw(double x) { return sin_or_cos_cw<false>(x); }
__device__ __forceinline__ double cos_cw(double x) { return sin_or_cos_cw<true>(x); }

// sin and cos of a small angle by short Taylor sums; anything larger goes through the general
// routine. LEVEL = 6: |d| < 2^-6, three coefficients each (truncation: sin d^9/9! < 1e-20
// relative, cos d^8/8! < 1e-19); LEVEL = 9: |d| < 2^-9, two coefficients each (sin d^7/7! and
// cos d^6/6! < 1e-19). Used for the drag correction of the mean anomaly and for the short-period
// corrections of the argument of latitude and the inclination (|d| < 1e-3 for every physical
// orbit), and for Newton corrections of Kepler's equation.
// The series is evaluated unconditionally (the common case, kept outside any branch so that
// it stays on the warp-uniform path); a lane whose angle is too large for it -- garbage
// states only -- overwrites the result through the general routine. The decision is per lane,
// so a result never depends on which other times share the warp.
// Taylor sums with NS sine coefficients (d^3 .. d^(2 NS + 1)) and NC cosine coefficients
// (d^2 .. d^(2 NC)); truncation errors |d|^(2 NS + 2) / (2 NS + 3)! relative (sin) and
// |d|^(2 NC + 2) / (2 NC + 2)! absolute (cos).
template <int NS, int NC>
__device__ __forceinline__ void sincos_taylor(double d, double &s, double &c) {
    static_assert(NS >= 2 && NS <= 4 && NC >= 2 && NC <= 5, "coefficients in the table");
    const double z = __dmul_rn(d, d);
    double p = kMc[MC_T3 + NS - 1];
#pragma unroll
    for (int i = NS - 2; i >= 0; --i) p = __fma_rn(p, z, kMc[MC_T3 + i]);
    double k = kMc[MC_U2 + NC - 1];
#pragma unroll
    for (int i = NC - 2; i >= 0; --i) k = __fma_rn(k, z, kMc[MC_U2 + i]);
    s = __fma_rn(__dmul_rn(d, z), p, d);
    c = __fma_rn(k, z, kMc[MC_ONE]);
}

template <int LEVEL>
__device__ __forceinline__ void sincos_series(double d, double &s, double &c) {
    if (LEVEL <= 6) {
        sincos_taylor<3, 3>(d, s, c);
    } else {
        sincos_taylor<2, 2>(d, s, c);
    }
}

// a b + c d and a b - c d with the contraction spelled out: left to the compiler, which of the
// two products is fused depends on the code around the expression, and the kernel's output
// modes (planes, records, mean elements, summary) would then differ from one another in the
// last bit.
__device__ __forceinline__ double mad2(double a, double b, double c, double d) {
    return __fma_rn(a, b, __dmul_rn(c, d));
}
__device__ __forceinline__ double msb2(double a, double b, double c, double d) {
    return __fma_rn(a, b, -__dmul_rn(c, d));
}

// (sin, cos) of a + d from those of a and of d
__device__ __forceinline__ void rotate(double sa, double ca, double sd, double cd, double &s,
                                       double &c) {
    s = __fma_rn(ca, sd, __dmul_rn(sa, cd));
    c = __fma_rn(-sa, sd, __dmul_rn(ca, cd));
}

template <int LEVEL = 6>
__device__ __forceinline__ void sincos_small(double d, double &s, double &c) {
    sincos_series<LEVEL>(d, s, c);
    if (!(fabs(d) < (LEVEL <= 6 ? 0.015625 : 0.001953125))) {
        sincos_cw(d, s, c);
    }
}

// atan(t) for |t| <= 0.25, 1 ulp (no reduction, no quadrant logic): the angle by which the
// Lyddane branch of dpper moves the node.
__device__ __forceinline__ double atan_small(double t) {
    const double z = __dmul_rn(t, t);
    double p = kMc[MC_A7];
#pragma unroll
    for (int i = 6; i >= 0; --i) p = __fma_rn(p, z, kMc[MC_A0 + i]);
    return __fma_rn(__dmul_rn(t, z), p, t);
}

// (1 + x)^(-2/3) for |x| < 2^-8: binomial series through x^6 (next term 0.38 x^7 < 6e-18).
__device__ __forceinline__ double pow_m2o3_1p(double x) {
    double p = kMc[MC_B6];
#pragma unroll
    for (int i = 4; i >= 0; --i) p = __fma_rn(p, x, kMc[MC_B1 + i]);
    return __fma_rn(p, x, kMc[MC_ONE]);
}

// pow(x, 2/3) for x > 0 (the reference passes x2o3 = 2.0/3.0, src/sgp4.cpp:1795, 1867)
__device__ __forceinline__ double pow_2o3(double x) {
    const double c = cbrt(x);
    return c * c;
}

}  // namespace pb

#endif
