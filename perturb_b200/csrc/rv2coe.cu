// N4: classical orbital elements from TEME states, batched -- the
// counterpart of perturb::ClassicalOrbitalElements(StateVector, GravModel)
// (src/perturb.cpp:119-131) and sgp4::rv2coe_SGP4 / newtonnu_SGP4 / angle_SGP4
// (src/sgp4.cpp:2863-3064, 2749-2803, 2659-2681), which the reference's verification
// printout applies to every propagated state (tests/test_perturb.cpp:385-397).
//
// Element-wise over (satellite, time): 48 B read + 88 B written per state, so the kernel's
// roofline is HBM (6.5 TB/s / 136 B = 4.8e10 states/s). Compiled with -fmad=false: the general
// routine rounds every product and sum on its own, as the reference's scalar code does, which
// keeps it on the reference's side of the 1e-8 "small" thresholds that switch elements to the
// 999999.1 "undefined" sentinel; the fast path for inclined ellipses spells its FMAs out.
#include <cuda_runtime.h>

#include "../../include/perturb_gpu.h"
#include "batch_internal.h"
#include "sgp4_math.cuh"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kSmall = 0.00000001;
constexpr double kUndefined = 999999.1;
constexpr double kInfinite = 999999.9;

// ---- general routine ------------------------------------------------------------------------
// Every conic and every degenerate case of rv2coe_SGP4 (src/sgp4.cpp:2863-3064), including the
// elements it reports as "undefined" (999999.1) or "infinite" (999999.9). Organised around
// what decides the outcome rather than around the reference's statement order:
//   * a Frame holds everything derived from (r, v) before any angle is taken;
//   * the shape of the orbit is two independent bits (circular, equatorial) -- the reference's
//     `typeorbit` 1..4 is their four combinations -- and each element states which
//     combinations it exists for;
//   * every angle comes from ONE helper, arc(): arccosine of a clamped cosine, mirrored to
//     (pi, 2 pi) by a sign test, or the sentinel when its vectors are too short to define it;
//   * true -> mean anomaly is one function of the conic type (newtonnu_SGP4, :2749-2803).
// It runs for the vanishing share of states the fast path below declines, as a real call.

__device__ __forceinline__ double dot3f(const double a[3], const double b[3]) {
    return __fma_rn(a[2], b[2], __fma_rn(a[1], b[1], __dmul_rn(a[0], b[0])));
}

struct Frame {
    double r_len, v_len, rdotv;
    double h[3], h_len;    // angular momentum r x v
    double node[3], node_len;  // line of nodes (-h1, h0, 0)
    double e[3], ecc;      // eccentricity vector
    double energy;         // specific mechanical energy v^2 / 2 - mu / r
};

__device__ inline Frame frame_of(const double r[3], const double v[3], double mus) {
    Frame f;
    f.r_len = sqrt(dot3f(r, r));
    f.v_len = sqrt(dot3f(v, v));
    f.rdotv = dot3f(r, v);
    f.h[0] = r[1] * v[2] - r[2] * v[1];
    f.h[1] = r[2] * v[0] - r[0] * v[2];
    f.h[2] = r[0] * v[1] - r[1] * v[0];
    f.h_len = sqrt(dot3f(f.h, f.h));
    f.node[0] = -f.h[1];
    f.node[1] = f.h[0];
    f.node[2] = 0.0;
    f.node_len = sqrt(dot3f(f.node, f.node));
    const double v2 = f.v_len * f.v_len, mu_over_r = mus / f.r_len;
    const double along_r = v2 - mu_over_r;
    for (int i = 0; i < 3; ++i) f.e[i] = (along_r * r[i] - f.rdotv * v[i]) / mus;
    f.ecc = sqrt(dot3f(f.e, f.e));
    f.energy = 0.5 * v2 - mu_over_r;
    return f;
}

// The angle in [0, pi] whose cosine is `cosine` (clamped: rounding can leave it a hair outside
// [-1, 1]), mirrored to 2 pi - angle when `mirror`; the "undefined" sentinel when `defined` is
// false. The reference mirrors angle_SGP4's results even when they are the sentinel
// (:2972-2974 and the like), so the mirror is applied after the select.
__device__ inline double arc(double cosine, bool defined, bool mirror) {
    const double c = fmin(fmax(cosine, -1.0), 1.0);
    const double a = defined ? acos(c) : kUndefined;
    return mirror ? kTwoPi - a : a;
}

// Angle between two vectors (angle_SGP4, :2659-2681): defined when neither is shorter than
// "small" in the product sense the reference uses.
__device__ inline double arc_between(const double a[3], double a_len, const double b[3],
                                     double b_len, bool mirror) {
    const double scale = a_len * b_len;
    return arc(dot3f(a, b) / scale, scale > kSmall * kSmall, mirror);
}

// Mean anomaly from true anomaly for any conic (newtonnu_SGP4; its eccentric-anomaly output is
// not used by rv2coe). "Infinite" where the reference leaves it so.
__device__ inline double mean_from_true(double ecc, double nu) {
    double m = kInfinite;
    const bool circular = fabs(ecc) < kSmall;
    const bool elliptic = !circular && ecc < 1.0 - kSmall;
    const bool hyperbolic = !circular && !elliptic && ecc > 1.0 + kSmall;
    if (circular) {
        m = nu;
    } else if (elliptic) {
        double sn, cn;
        sincos(nu, &sn, &cn);
        const double inv = 1.0 / (1.0 + ecc * cn);
        const double sin_e = sqrt(1.0 - ecc * ecc) * sn * inv, cos_e = (ecc + cn) * inv;
        const double big_e = atan2(sin_e, cos_e);
        m = big_e - ecc * sin(big_e);
    } else if (hyperbolic) {
        if (fabs(nu) + 0.00001 < kPi - acos(1.0 / ecc)) {
            const double sinh_f = sqrt(ecc * ecc - 1.0) * sin(nu) / (1.0 + ecc * cos(nu));
            const double f = log(sinh_f + sqrt(sinh_f * sinh_f + 1.0));  // asinh, :2705-2711
            m = ecc * sinh(f) - f;
        }
    } else if (fabs(nu) < 168.0 * kPi / 180.0) {  // parabolic band
        const double d = tan(0.5 * nu);
        m = d + d * d * d / 3.0;
    }
    if (ecc < 1.0) {  // :2796-2801, applied to whatever m holds
        m = fmod(m, kTwoPi);
        if (m < 0.0) m += kTwoPi;
    }
    return m;
}

// out[0..10] = p, a, ecc, incl, omega, argp, nu, m, arglat, truelon, lonper.
__device__ inline void rv2coe(const double r[3], const double v[3], double mus, double out[11]) {
    const Frame f = frame_of(r, v, mus);
    if (!(f.h_len > kSmall)) {  // rectilinear or degenerate: nothing is defined (:3049-3062)
        for (int c = 0; c < 11; ++c) out[c] = kUndefined;
        return;
    }
    const double incl = acos(f.h[2] / f.h_len);
    const bool circular = f.ecc < kSmall;
    const bool equatorial = (incl < kSmall) || (fabs(incl - kPi) < kSmall);
    const bool retrograde = incl > 0.5 * kPi;
    // which elements exist for which shape (the reference's typeorbit: 1 = neither bit,
    // 2 = both, 3 = circular only, 4 = equatorial only)
    const bool has_argp = !circular && !equatorial;
    const bool has_nu = !circular;
    const bool has_arglat = circular && !equatorial;
    const bool has_truelon = circular && equatorial && f.r_len > kSmall;
    const bool has_lonper = !circular && equatorial && f.ecc > kSmall;

    out[0] = f.h_len * f.h_len / mus;
    out[1] = (fabs(f.energy) > kSmall) ? -mus / (2.0 * f.energy) : kInfinite;
    out[2] = f.ecc;
    out[3] = incl;
    out[4] = arc(f.node[0] / f.node_len, f.node_len > kSmall, f.node_len > kSmall && f.node[1] < 0.0);
    out[5] = has_argp ? arc_between(f.node, f.node_len, f.e, f.ecc, f.e[2] < 0.0) : kUndefined;
    const double nu = has_nu ? arc_between(f.e, f.ecc, r, f.r_len, f.rdotv < 0.0) : kUndefined;
    out[6] = nu;
    const double arglat =
        has_arglat ? arc_between(f.node, f.node_len, r, f.r_len, r[2] < 0.0) : kUndefined;
    out[8] = arglat;
    // the two longitudes are measured from the x axis and counted backwards on retrograde orbits
    double truelon = kUndefined, lonper = kUndefined;
    if (has_truelon) {
        truelon = arc(r[0] / f.r_len, true, r[1] < 0.0);
        if (retrograde) truelon = kTwoPi - truelon;
    }
    if (has_lonper) {
        lonper = arc(f.e[0] / f.ecc, true, f.e[1] < 0.0);
        if (retrograde) lonper = kTwoPi - lonper;
    }
    out[9] = truelon;
    out[10] = lonper;
    // mean anomaly: from the true anomaly where there is one, else the angle that replaces it
    out[7] = has_nu ? mean_from_true(f.ecc, nu) : has_arglat ? arglat : has_truelon ? truelon : 0.0;
}

// ---- fast path ----------------------------------------------------------------------------
// With libdevice's acos / atan2 / sin / sqrt and IEEE divisions the routine above costs about
// as many instructions as an SGP4 evaluation and ran at 0.39 of the HBM roofline (136 B per
// state). Every state SGP4 produces for a catalogued object is an inclined ellipse ("typeorbit
// 1" with e < 1), so that case gets its own straight-line code: reciprocal square roots
// instead of sqrt + divide, cos(nu) and sin(nu) from the dot product the angle came from, sin E
// from the unit vector handed to atan2, and polynomial arccosine / arctangent without slow
// paths (scripts/fit_asin.py, scripts/fit_atan.py: 1 ulp kernels). Anything else -- circular,
// equatorial, parabolic, hyperbolic, degenerate -- takes the routine above unchanged, so the
// "undefined" / "infinite" sentinels are exactly the reference's.
__constant__ double kAsinP[12] = {
    1.66666666666666490881e-01, 7.50000000002076366856e-02, 4.46428571034236457149e-02,
    3.03819473670984795122e-02, 2.23720476317445099435e-02, 1.73552599557863229673e-02,
    1.39296529023266325159e-02, 1.18754943826369224746e-02, 7.80294947735331746036e-03,
    1.60355143491488216778e-02, -1.07490503396978076545e-02, 2.81692180608814138343e-02,
};
__constant__ double kAtanP[11] = {
    -3.33333333333333314830e-01, 1.99999999999955047070e-01, -1.42857142846629397992e-01,
    1.11111110149666403313e-01, -9.09090456635439914113e-02, 7.69218291746116139551e-02,
    -6.66450759756167521175e-02, 5.85811547629988410568e-02, -5.08527480164209441660e-02,
    3.92266055108713568300e-02, -1.91707003937429884544e-02,
};
constexpr double kPio2Hi = 1.57079632679489655800e+00, kPio2Lo = 6.12323399573676603587e-17;
constexpr double kPiLo = 1.22464679914735320717e-16;  // pi - (double)pi
constexpr double kPio4Hi = 7.85398163397448278999e-01, kPio4Lo = 3.06161699786838301793e-17;

// acos(x) for |x| <= 1: pi/2 - asin(x) below 0.5, 2 asin(sqrt((1 - |x|) / 2)) above (1 - |x| is
// exact there), asin(t) = t + t^3 P(t^2) on t^2 <= 0.25.
__device__ __forceinline__ double acos_fast(double x) {
    const double ax = fabs(x);
    const bool big = ax > 0.5;
    const double zb = __fma_rn(-0.5, ax, 0.5);
    const double z = big ? zb : __dmul_rn(x, x);
    const double t = big ? pb::sqrt_fast(zb) : x;
    double p = kAsinP[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = __fma_rn(p, z, kAsinP[i]);
    const double q = __fma_rn(__dmul_rn(t, z), p, t);  // asin(t)
    const double small = __dsub_rn(kPio2Hi, __dsub_rn(q, kPio2Lo));
    const double twice = __dadd_rn(q, q);
    const double large = (x < 0.0) ? __dsub_rn(kPi, __dsub_rn(twice, kPiLo)) : twice;
    return big ? large : small;
}

// atan2(y, x) for finite arguments, not both zero: one division (the reduced tangent
// (mn - mx) / (mn + mx) above tan(pi/8), mn / mx below), atan(t) = t + t^3 P(t^2) on
// |t| <= 0.4143, then the octant is unfolded.
__device__ __forceinline__ double atan2_fast(double y, double x) {
    const double ay = fabs(y), ax = fabs(x);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const bool upper = mn > __dmul_rn(0.41421356237309503, mx);
    const double num = upper ? __dsub_rn(mn, mx) : mn;
    const double den = upper ? __dadd_rn(mn, mx) : mx;
    const double t = pb::div_fast(num, den);
    const double z = __dmul_rn(t, t);
    double p = kAtanP[10];
#pragma unroll
    for (int i = 9; i >= 0; --i) p = __fma_rn(p, z, kAtanP[i]);
    double r = __fma_rn(__dmul_rn(t, z), p, t);
    if (upper) r = __dadd_rn(kPio4Hi, __dadd_rn(r, kPio4Lo));
    if (ay > ax) r = __dsub_rn(kPio2Hi, __dsub_rn(r, kPio2Lo));
    if (x < 0.0) r = __dsub_rn(kPi, __dsub_rn(r, kPiLo));
    return (y < 0.0) ? -r : r;
}

__device__ __forceinline__ double clamp1(double c) { return fmin(fmax(c, -1.0), 1.0); }

// Returns false (nothing written) unless the state is an inclined ellipse well inside every
// threshold of rv2coe_SGP4; the caller then runs the general routine.
__device__ __forceinline__ bool rv2coe_fast(const double r[3], const double v[3], double mus,
                                            double inv_mus, double out[11]) {
    const double rr = dot3f(r, r), vv = dot3f(v, v), rdotv = dot3f(r, v);
    const double h[3] = {__fma_rn(r[1], v[2], -__dmul_rn(r[2], v[1])),
                         __fma_rn(r[2], v[0], -__dmul_rn(r[0], v[2])),
                         __fma_rn(r[0], v[1], -__dmul_rn(r[1], v[0]))};
    const double hh = dot3f(h, h);
    const double nn = __fma_rn(h[1], h[1], __dmul_rn(h[0], h[0]));  // nbar = (-h1, h0, 0)
    const double inv_r = pb::rsqrt_fast(rr), inv_h = pb::rsqrt_fast(hh), inv_n = pb::rsqrt_fast(nn);
    const double mu_r = __dmul_rn(mus, inv_r);
    const double c1 = __dsub_rn(vv, mu_r);
    double e[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        e[i] = __dmul_rn(__fma_rn(c1, r[i], -__dmul_rn(rdotv, v[i])), inv_mus);
    }
    const double ee = dot3f(e, e);
    const double inv_e = pb::rsqrt_fast(ee);
    const double ecc = __dmul_rn(ee, inv_e);
    const double sme = __fma_rn(0.5, vv, -mu_r);
    const double hk = __dmul_rn(h[2], inv_h);
    // margins: far from magh, magn, ecc ~ 1e-8, ecc ~ 1, sme ~ 0 and from the equatorial cut
    // (|hk| = cos 1e-8 is 1 - 5e-17: anything below 1 - 1e-12 is safely inclined)
    const bool regular = (hh > 1.0e-12) && (nn > 1.0e-12) && (ecc > 1.0e-6) && (ecc < 0.999999) &&
                         (fabs(hk) < 0.999999999999) && (fabs(sme) > 1.0e-6) && (rr > 1.0e-12);
    // (no early return: with one, everything below counts as divergent code and loses the
    // uniform datapath for its constants; an irregular state just computes garbage here)

    const double incl = acos_fast(hk);
    double omega = acos_fast(clamp1(__dmul_rn(-h[1], inv_n)));
    if (h[0] < 0.0) omega = __dsub_rn(kTwoPi, omega);
    const double ne = __fma_rn(h[0], e[1], -__dmul_rn(h[1], e[0]));  // nbar . ebar
    double argp = acos_fast(clamp1(__dmul_rn(__dmul_rn(ne, inv_n), inv_e)));
    if (e[2] < 0.0) argp = __dsub_rn(kTwoPi, argp);
    const double cnu = clamp1(__dmul_rn(__dmul_rn(dot3f(e, r), inv_e), inv_r));
    double nu = acos_fast(cnu);
    if (rdotv < 0.0) nu = __dsub_rn(kTwoPi, nu);
    // sin(nu) = +-sqrt((1 - c)(1 + c)): what sin(acos(c)) is, without the round trip
    double snu = pb::sqrt_fast(__dmul_rn(__dsub_rn(1.0, cnu), __dadd_rn(1.0, cnu)));
    if (rdotv < 0.0) snu = -snu;
    // newtonnu_SGP4, elliptical branch (src/sgp4.cpp:2768-2775)
    const double rden = pb::rcp_fast(__fma_rn(ecc, cnu, 1.0));
    const double beta = pb::sqrt_fast(__fma_rn(-ecc, ecc, 1.0));
    const double sine = __dmul_rn(__dmul_rn(beta, snu), rden);
    const double cose = __dmul_rn(__dadd_rn(ecc, cnu), rden);
    const double e0 = atan2_fast(sine, cose);
    // (sine, cose) is a unit vector by an identity in nu, so sin(e0) is `sine`
    double m = __fma_rn(-ecc, sine, e0);
    m = pb::fmod_2pi(m);
    if (m < 0.0) m = __dadd_rn(m, kTwoPi);

    out[0] = __dmul_rn(hh, inv_mus);
    out[1] = __dmul_rn(-0.5 * mus, pb::rcp_fast(sme));
    out[2] = ecc;
    out[3] = incl;
    out[4] = omega;
    out[5] = argp;
    out[6] = nu;
    out[7] = m;
    out[8] = kUndefined;
    out[9] = kUndefined;
    out[10] = kUndefined;
    return regular;
}

// The general routine as a real call that stores its own results: it runs for a vanishing share
// of the states, and inlined it would set the register allocation (94) of the whole kernel.
__device__ __noinline__ void rv2coe_general_store(double rx, double ry, double rz, double vx,
                                                  double vy, double vz, double mus, double *dst,
                                                  size_t plane_stride) {
    double r[3] = {rx, ry, rz}, v[3] = {vx, vy, vz}, out[11];
    rv2coe(r, v, mus, out);
#pragma unroll
    for (int c = 0; c < 11; ++c) dst[c * plane_stride] = out[c];
}

// grid = (time blocks, satellites [, satellites / 65535]): no index division per thread
__global__ void __launch_bounds__(256, 4)
rv2coe_kernel(const double *__restrict__ posvel, size_t n, size_t m, size_t row_stride,
              size_t plane_stride, double mus, double inv_mus, double *__restrict__ coe,
              size_t coe_plane_stride) {
    const size_t i_own = static_cast<size_t>(blockIdx.z) * 65535u + blockIdx.y;
    const size_t k_own = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    // lanes past the end of the row (or rows past the last satellite, in the last grid slab)
    // redo a state that exists and store nothing: no early return, so the arithmetic below
    // stays warp-uniform code
    const bool in_row = (k_own < m) && (i_own < n);
    const size_t i = i_own < n ? i_own : n - 1;
    const size_t at = i * row_stride + (k_own < m ? k_own : m - 1);  // lanes = consecutive times
    double r[3], v[3], out[11];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = posvel[c * plane_stride + at];
        v[c] = posvel[(c + 3) * plane_stride + at];
    }
    const bool have = !(isnan(r[0]) || isnan(v[0]));  // codes 1-4 carry no state
    const bool fast = rv2coe_fast(r, v, mus, inv_mus, out);
    if (in_row) {
        if (have && !fast) {
            rv2coe_general_store(r[0], r[1], r[2], v[0], v[1], v[2], mus, coe + at, coe_plane_stride);
        } else {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int c = 0; c < 11; ++c) coe[c * coe_plane_stride + at] = have ? out[c] : nan;
        }
    }
}

// Device that owns `p` (device or managed memory), or -1.
int owner_of(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? attr.device
                                                                                     : -1;
}

struct OnDevice {  // make `dev` current for the scope of a call
    int prev;
    bool ok;
    explicit OnDevice(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~OnDevice() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

extern "C" int perturb_gpu_rv2coe(const double *posvel, size_t n, size_t m, size_t row_stride,
                                  size_t plane_stride, int grav_model, double *coe,
                                  size_t coe_plane_stride, void *stream) {
    if (posvel == nullptr || coe == nullptr || row_stride < m || plane_stride < n * row_stride ||
        coe_plane_stride < n * row_stride || (m + 255) / 256 > 2147483647ULL ||
        n > 65535ULL * 65535ULL) {
        return pb::report_failure(PERTURB_GPU_ERR_ARG, "perturb_gpu_rv2coe: null pointer or inconsistent strides");
    }
    double mus;  // getgravconst, src/sgp4.cpp:2119,2130,2141
    switch (grav_model) {
        case PERTURB_GPU_WGS72_OLD: mus = 398600.79964; break;
        case PERTURB_GPU_WGS72: mus = 398600.8; break;
        case PERTURB_GPU_WGS84: mus = 398600.5; break;
        default: return pb::report_failure(PERTURB_GPU_ERR_ARG, "perturb_gpu_rv2coe: unknown gravity model");
    }
    if (n == 0 || m == 0) return PERTURB_GPU_OK;
    // both buffers in the memory of ONE device: the kernel runs there, whatever device is
    // current in the calling thread
    const int dev = owner_of(posvel);
    if (dev < 0 || owner_of(coe) != dev) {
        return pb::report_failure(PERTURB_GPU_ERR_ARG,
                                  "perturb_gpu_rv2coe: posvel and coe must be memory of the same device");
    }
    const OnDevice guard(dev);
    if (!guard.ok) return pb::report_failure(PERTURB_GPU_ERR_CUDA, "perturb_gpu_rv2coe: cudaSetDevice failed");
    // CTA width that wastes the fewest lanes at the end of a row (1440 times: 9 x 160, none)
    unsigned threads = 256;
    size_t waste = ((m + 255) / 256) * 256 - m;
    for (unsigned t = 224; t >= 128; t -= 32) {
        const size_t w = ((m + t - 1) / t) * t - m;
        if (w < waste) {
            waste = w;
            threads = t;
        }
    }
    const dim3 grid(static_cast<unsigned>((m + threads - 1) / threads),
                    static_cast<unsigned>(n < 65535 ? n : 65535),
                    static_cast<unsigned>((n + 65534) / 65535));
    rv2coe_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        posvel, n, m, row_stride, plane_stride, mus, 1.0 / mus, coe, coe_plane_stride);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        return pb::report_failure(PERTURB_GPU_ERR_CUDA, cudaGetErrorString(e));
    }
    return PERTURB_GPU_OK;
}
