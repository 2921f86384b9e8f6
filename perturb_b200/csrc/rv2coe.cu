// N4: classical orbital elements from TEME states, batched -- the
// counterpart of perturb::ClassicalOrbitalElements(StateVector, GravModel)
// (src/perturb.cpp:119-131) and sgp4::rv2coe_SGP4 / newtonnu_SGP4 / angle_SGP4
// (src/sgp4.cpp:2863-3064, 2749-2803, 2659-2681), which the reference's verification
// printout applies to every propagated state (tests/test_perturb.cpp:385-397).
//
// Element-wise over (satellite, time): 48 B read + 88 B written per state, so the kernel's
// roofline is HBM (6.5 TB/s / 136 B = 4.8e10 states/s). Compiled with -fmad=false: the general
// routine's arithmetic is the reference's apart from libdevice-vs-glibc acos/atan2/sin/cos,
// which matters for the 1e-8 "small" thresholds that switch outputs to the 999999.1
// "undefined" sentinel; the fast path for inclined ellipses (below) spells its FMAs out.
#include <cuda_runtime.h>

#include "../../include/perturb_gpu.h"
#include "sgp4_math.cuh"

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

__device__ __forceinline__ double dot3f(const double a[3], const double b[3]) {
    return __fma_rn(a[2], b[2], __fma_rn(a[1], b[1], __dmul_rn(a[0], b[0])));
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
        coe_plane_stride < n * row_stride || (m + 255) / 256 > 2147483647ULL ||
        n > 65535ULL * 65535ULL) {
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
    return cudaGetLastError() == cudaSuccess ? PERTURB_GPU_OK : PERTURB_GPU_ERR_CUDA;
}
