// One SGP4/SDP4 evaluation: the device-side counterpart of
//   sgp4::sgp4()  src/sgp4.cpp:1769-2065   (secular + drag, Kepler solve, short-period, TEME)
//   dspace()      src/sgp4.cpp:1018-1166   (deep-space secular rates, resonance integrator)
//   dpper('n')    src/sgp4.cpp:235-368     (lunar-solar periodics, Lyddane modification)
// written for the propagation layout of sgp4_fields.h: per-satellite constants come from a
// provider `K` (shared-memory tile in kernel 2, registers in kernel 1's epoch probe) and the
// orbit class is a template parameter so each specialisation carries only its own terms.
//
// Differences from the reference, all value-preserving up to rounding:
//   * stateless: the resonance integrator always starts from epoch (the reference's
//     atime/xli/xni resume cache is an optimisation; restarted and resumed walks are
//     bit-identical);
//   * near-Earth: pow(xke/nm, 2/3) is per-satellite constant (nm == no_unkozai) and comes
//     from PF_AOBASE; sin/cos(inclo) come from PF_SINIO/PF_COSIO;
//   * bstar*cc4 and bstar*cc5 are hoisted (the reference evaluates them left to right, so
//     the hoisted products are the same doubles);
//   * libm calls are replaced by the routines of sgp4_math.cuh, divisions by reciprocal-
//     multiply, atan2 + sincos of the argument of latitude by a rotation of the (unit) vector
//     (sinu, cosu), sqrt(pl) by sqrt(am) sqrt(1 - el2), and sin/cos of angles that differ from
//     an angle with known sin/cos by a small amount (drag correction of M, the solar / lunar
//     equation of centre in dpper, the perturbed inclination) by short Taylor rotations;
//   * Kepler's equation: closed form for nearly circular orbits, Halley's iteration for
//     e^2 < 0.64, the reference's Newton loop for the rest (kepler_small, kepler_halley,
//     kepler_newton): all land on the root the reference's loop converges to;
//   * near-Earth full-drag class: sin/cos of the argument of latitude by rotating sin/cos(xmdf)
//     through argpp and the small drag / long-period angle, without the reference's three
//     fmod reductions (kRotU in sgp4_eval);
//   * deep space: sin/cos of the solar / lunar mean anomalies from a per-satellite and a
//     per-time part that kernel 2 evaluates once per CTA (ThirdBodyPhase).
// The record fields the reference overwrites on every call (t, error, am..nm, and for deep
// space aycof/xlcof/con41/x1mth2/x7thm1) are returned through `Sgp4Side` when asked for.
#ifndef PERTURB_B200_SGP4_EVAL_CUH
#define PERTURB_B200_SGP4_EVAL_CUH

#include "sgp4_fields.h"
#include "sgp4_math.cuh"

#ifndef PB_KEPLER_SMALL
#define PB_KEPLER_SMALL 1  // 0: every orbit through the Newton iteration (bisecting / comparison builds)
#endif

#ifndef PB_KEPLER_HALLEY
#define PB_KEPLER_HALLEY 1  // 0: eccentric orbits through the reference's Newton iteration only
#endif
#ifndef PB_KEPLER_ROT
#define PB_KEPLER_ROT 1  // 0: sin/cos of the argument of latitude from its reduced value (comparison builds)
#endif

namespace pb {

// getgravconst() results the evaluation needs (src/sgp4.cpp:2101-2157), batch-wide.
struct GravConsts {
    double radiusearthkm, xke, j2, j3oj2, vkmpersec, inv_xke;
};

struct ClassFlags {  // used when CLS == PB_CLS_RUNTIME
    bool deep, isimp;
    int irez;
};

struct Sgp4Side {  // what sgp4() leaves in the record (src/sgp4.cpp:1895-1901, 1952-1957, 2017-2019)
    double am, em, im, Om, om, mm, nm;
    double aycof, xlcof, con41, x1mth2, x7thm1;
    int stage;  // 0: nothing stored, 1: mean elements, 2: + aycof/xlcof, 3: + con41..x7thm1
    // resonant deep-space satellites: the integrator state dspace() leaves behind
    // (src/sgp4.cpp:1143-1147), always stored (dspace runs before the first error check)
    double atime, xli, xni;
    bool resonant;
};

// dpper's solar and lunar mean anomalies, zm = zmos + zns t and zmol + znl t (src/sgp4.cpp:262,
// 286), advance at rates that are the same for every satellite. With t = t_time - t_sat (minutes
// of the evaluation time and of the satellite's epoch from one fixed reference) the angle splits
// into a per-satellite part (zmos - zns t_sat) and a per-time part (zns t_time): kernel 2 takes
// the general sin/cos of either part ONCE per CTA (per staged satellite, per staged time) and an
// evaluation only rotates one by the other -- 4 FP64 operations per body instead of a general
// sin/cos (19 + quadrant logic + 9 constant loads). The parts are functions of the satellite and
// of the time alone (fixed reference date, not the grid), so a result still does not depend on
// how a grid is tiled or chunked. |angle error| from the split: a few 1e-13 rad (the parts are
// ~2e3 rad), i.e. 1e-16 of the periodic terms.
struct ThirdBodyPhase {
    double ss, cs;  // sin, cos of the Sun's zm
    double sm, cm;  // sin, cos of the Moon's zm
};
constexpr double kPhaseRefJd = 2451545.0;
// minutes from the reference date: of a grid time (JD pair, or minutes from epoch as they are) ...
__device__ __forceinline__ double phase_minutes_of_time(int use_jd, double ta, double tb) {
    return use_jd ? __dmul_rn(__dadd_rn(__dsub_rn(ta, kPhaseRefJd), tb), 1440.0) : ta;
}
// ... and of a satellite's epoch (zero for the minutes-from-epoch entry point)
__device__ __forceinline__ double phase_minutes_of_epoch(int use_jd, double jd0, double jdf0) {
    return use_jd ? __dmul_rn(__dadd_rn(__dsub_rn(jd0, kPhaseRefJd), jdf0), 1440.0) : 0.0;
}
__device__ __forceinline__ void phase_of_time(double minutes, double2 &sun, double2 &moon) {
    sincos_cw(__dmul_rn(kMc[MC_ZNS], minutes), sun.x, sun.y);
    sincos_cw(__dmul_rn(kMc[MC_ZNL], minutes), moon.x, moon.y);
}
__device__ __forceinline__ void phase_of_satellite(double zmos, double zmol, double epoch_minutes,
                                                   double2 &sun, double2 &moon) {
    sincos_cw(__fma_rn(-kMc[MC_ZNS], epoch_minutes, zmos), sun.x, sun.y);
    sincos_cw(__fma_rn(-kMc[MC_ZNL], epoch_minutes, zmol), moon.x, moon.y);
}
__device__ __forceinline__ ThirdBodyPhase phase_combine(double2 sat_sun, double2 sat_moon,
                                                        double2 time_sun, double2 time_moon) {
    ThirdBodyPhase ph;
    rotate(sat_sun.x, sat_sun.y, time_sun.x, time_sun.y, ph.ss, ph.cs);
    rotate(sat_moon.x, sat_moon.y, time_moon.x, time_moon.y, ph.sm, ph.cm);
    return ph;
}

template <int CLS>
struct ClassTraits {
    static constexpr bool kStatic = true;
    static constexpr bool kDeep = (CLS >= PB_CLS_DEEP_NR);
    static constexpr bool kSimp = (CLS != PB_CLS_NEAR_FULL);
    static constexpr int kIrez =
        (CLS == PB_CLS_DEEP_R1) ? 1 : (CLS == PB_CLS_DEEP_R2) ? 2 : 0;
};
template <>
struct ClassTraits<PB_CLS_RUNTIME> {
    static constexpr bool kStatic = false;
    static constexpr bool kDeep = false, kSimp = false;
    static constexpr int kIrez = 0;
};

// Resonance terms at one integrator node, src/sgp4.cpp:1097-1126.
//
// The reference evaluates 3 (24 h) or 10 (12 h) sines and as many cosines of sums of xli, xomi
// and fixed phases. Here sin/cos of xli (and xomi) are taken ONCE and every term follows
// from angle-addition / double- / triple-angle identities with the phases' sines and cosines
// as compile-time constants: 2 sincos + ~70 FMA-class operations instead of 10 sincos with
// their range reductions (12 h case). Value-preserving up to rounding (a few 1e-16).
// Written with explicit fused operations: the node table (built by a pre-pass kernel) and the
// in-line fall-back must produce the same bits no matter how the compiler would contract
// either copy.
namespace res {
// sin / cos of the phase constants of src/sgp4.cpp:1036-1043 live in the constant table
// (kMc[MC_S_FASX2] ...): the doubles the reference uses, correctly rounded.

// (sin, cos) of a + b from those of a and b
__device__ __forceinline__ void add(double sa, double ca, double sb, double cb, double &s,
                                    double &c) {
    s = __fma_rn(sa, cb, __dmul_rn(ca, sb));
    c = __fma_rn(ca, cb, -__dmul_rn(sa, sb));
}
// (sin, cos) of a - b
__device__ __forceinline__ void sub(double sa, double ca, double sb, double cb, double &s,
                                    double &c) {
    s = __fma_rn(sa, cb, -__dmul_rn(ca, sb));
    c = __fma_rn(ca, cb, __dmul_rn(sa, sb));
}
// accumulate weight * (sin, cos)(a - phase) into (xs, xc), for two terms sharing the phase:
// w1 (S1, C1) + w2 (S2, C2), then one rotation by the phase
__device__ __forceinline__ void pair(double w1, double s1, double c1, double w2, double s2,
                                     double c2, double sg, double cg, double &xs, double &xc) {
    const double ss = __fma_rn(w2, s2, __dmul_rn(w1, s1));
    const double cc = __fma_rn(w2, c2, __dmul_rn(w1, c1));
    xs = __dadd_rn(xs, __fma_rn(ss, cg, -__dmul_rn(cc, sg)));
    xc = __dadd_rn(xc, __fma_rn(cc, cg, __dmul_rn(ss, sg)));
}
}  // namespace res

template <class K>
__device__ __forceinline__ void resonance_rates(const K &k, int irez, double xli, double xni,
                                                double atime, double &xndt, double &xldot,
                                                double &xnddt) {
    double sl, cl;
    sincos_cw(xli, sl, cl);
    if (irez != 2) {
        // theta1 = xli - fasx2, theta2 = 2 (xli - fasx4), theta3 = 3 (xli - fasx6)
        double s1, c1, s, c;
        res::sub(sl, cl, kMc[MC_S_FASX2], kMc[MC_C_FASX2], s1, c1);
        res::sub(sl, cl, kMc[MC_S_FASX4], kMc[MC_C_FASX4], s, c);
        const double s2 = __dmul_rn(__dadd_rn(s, s), c);
        const double c2 = __fma_rn(-__dadd_rn(s, s), s, 1.0);
        res::sub(sl, cl, kMc[MC_S_FASX6], kMc[MC_C_FASX6], s, c);
        const double f = __fma_rn(-4.0 * s, s, 1.0);        // 1 - 4 s^2
        const double s3 = __dmul_rn(s, __dadd_rn(f, 2.0));  // s (3 - 4 s^2)
        const double c3 = __dmul_rn(c, f);                   // c (4 c^2 - 3) = c (1 - 4 s^2)
        const double del1 = k(PF_DEL1), del2 = k(PF_DEL2), del3 = k(PF_DEL3);
        xndt = __fma_rn(del3, s3, __fma_rn(del2, s2, __dmul_rn(del1, s1)));
        xldot = __dadd_rn(xni, k(PF_XFACT));
        xnddt = __fma_rn(__dmul_rn(3.0, del3), c3,
                         __fma_rn(__dmul_rn(2.0, del2), c2, __dmul_rn(del1, c1)));
        xnddt = __dmul_rn(xnddt, xldot);
    } else {
        const double xomi = __fma_rn(k(PF_ARGPDOT), atime, k(PF_ARGPO));
        double so, co;
        sincos_cw(xomi, so, co);
        const double s2o = __dmul_rn(__dadd_rn(so, so), co), c2o = __fma_rn(-__dadd_rn(so, so), so, 1.0);
        const double s2l = __dmul_rn(__dadd_rn(sl, sl), cl), c2l = __fma_rn(-__dadd_rn(sl, sl), sl, 1.0);
        double sA, cA, sB, cB;
        double xs = 0.0, xc = 0.0;   // sum of d sin(..), sum of d cos(..) for the l = 2 terms
        double ys = 0.0, yc = 0.0;   // same for the terms the reference doubles in xnddt (44, 54)
        // g22: (2 xomi + xli), (xli)
        res::add(s2o, c2o, sl, cl, sA, cA);
        res::pair(k(PF_D2201), sA, cA, k(PF_D2211), sl, cl, kMc[MC_S_G22], kMc[MC_C_G22], xs, xc);
        // g32 and g52 share (xomi + xli), (xli - xomi)
        res::add(so, co, sl, cl, sA, cA);
        res::sub(sl, cl, so, co, sB, cB);
        res::pair(k(PF_D3210), sA, cA, k(PF_D3222), sB, cB, kMc[MC_S_G32], kMc[MC_C_G32], xs, xc);
        res::pair(k(PF_D5220), sA, cA, k(PF_D5232), sB, cB, kMc[MC_S_G52], kMc[MC_C_G52], xs, xc);
        // g44: (2 xomi + 2 xli), (2 xli)
        res::add(s2o, c2o, s2l, c2l, sA, cA);
        res::pair(k(PF_D4410), sA, cA, k(PF_D4422), s2l, c2l, kMc[MC_S_G44], kMc[MC_C_G44], ys, yc);
        // g54: (xomi + 2 xli), (2 xli - xomi)
        res::add(so, co, s2l, c2l, sA, cA);
        res::sub(s2l, c2l, so, co, sB, cB);
        res::pair(k(PF_D5421), sA, cA, k(PF_D5433), sB, cB, kMc[MC_S_G54], kMc[MC_C_G54], ys, yc);
        xndt = __dadd_rn(xs, ys);
        xldot = __dadd_rn(xni, k(PF_XFACT));
        xnddt = __dmul_rn(__fma_rn(2.0, yc, xc), xldot);
    }
}

// One +-720 min Euler-Maclaurin step of the resonance integrator, src/sgp4.cpp:1143-1145.
__device__ __forceinline__ void resonance_advance(double xndt, double xldot, double xnddt,
                                                  double delt, double &xli, double &xni,
                                                  double &atime) {
    xli = __fma_rn(xndt, kMc[MC_STEP2], __fma_rn(xldot, delt, xli));
    xni = __fma_rn(xnddt, kMc[MC_STEP2], __fma_rn(xndt, delt, xni));
    atime = __dadd_rn(atime, delt);
}

// Number of whole 720-minute steps the reference's loop `while (fabs(t - atime) >= 720)`
// takes from epoch (src/sgp4.cpp:1093-1147): floor(|t| / 720), decided on the exact
// remainder |t| - 720 n (exactly representable) so that the count is the reference's.
__device__ __forceinline__ int resonance_step_count(double t) {
    const double at = fabs(t);
    // (selects throughout: as branches this was a reconvergence region per evaluation)
    const bool sane = at < 1.0e9;  // non-finite or absurd time: garbage in, no spinning
    int n = static_cast<int>((sane ? at : 0.0) * (1.0 / 720.0));
    const double rem = __fma_rn(-720.0, static_cast<double>(n), at);
    n += (rem >= 720.0) ? 1 : ((rem < 0.0) ? -1 : 0);
    return sane ? n : 0;
}

// Integrator nodes tabulated per satellite by the pre-pass kernel: record j holds the state
// (xli, xni) after (k0 + j) steps from epoch (negative = backwards) and the rates (xndt, xnddt)
// the integrator derives from that state (src/sgp4.cpp:1097-1140). The rates depend on the node
// only, so an evaluation whose node is tabulated needs no sin/cos of the resonance angles at
// all: one 32-byte record (one sector) instead.
struct ResNode {
    double xli, xni, xndt, xnddt;
};
struct ResNodes {
    const ResNode *table;  // all records; this satellite's are [first, first + count)
    int first;             // record index of this satellite's node k0
    int k0, count;
};

// Kepler's equation as the reference solves it (src/sgp4.cpp:1965-1983): Newton iteration from
// E = u, at most 10 corrections tem5, each clamped to +-0.95, stop after one smaller than 1e-12;
// the sin/cos it keeps are those of the iterate BEFORE that last correction, which is never
// applied to them.
//
// Three observations make this cheaper without leaving the reference's fixed point:
//  * Newton converges quadratically, so once a correction d is below 6e-8 the NEXT one
//    would be < e d^2 / (2 (1 - e)) < 1e-12: the reference's next pass only re-evaluates
//    sin/cos at E + d and stops. That pass is replaced by rotating (sin E, cos E) by d
//    (second-order series, error < 1e-22) after the loop.
//  * a lane that has finished only stops moving its iterate: the later passes the rest of
//    the warp needs recompute the very same sin/cos and correction from the frozen
//    iterate (same instructions, same inputs), so nothing has to be saved or selected per
//    pass, and a result never depends on how many iterations the other lanes of the warp
//    (other times of the same satellite) need.
//  * the clamp is two selects; a NaN correction stays NaN and stops the lane, as in the
//    reference.
template <bool UNIFORM>
__device__ __forceinline__ void kepler_newton(double u, double axnl, double aynl, double &sineo1,
                                              double &coseo1) {
    double eo1 = u, tem5, mag;
    bool done = false;
    int ktr = 1;
#pragma unroll 1
    do {
        sincos_cw(eo1, sineo1, coseo1);
        tem5 = 1.0 - coseo1 * axnl - sineo1 * aynl;
        tem5 = div_coarse(u - aynl * coseo1 + axnl * sineo1 - eo1, tem5);
        tem5 = (tem5 > kMc[MC_0P95]) ? kMc[MC_0P95] : tem5;
        tem5 = (tem5 < -kMc[MC_0P95]) ? -kMc[MC_0P95] : tem5;
        mag = fabs(tem5);
        // stop as evaluated (< 1e-12), or stop with one confirming rotation owed (< 6e-8)
        const bool stop = !(mag >= kMc[MC_1EM12]) || (mag < 5.9604644775390625e-8 && ktr < 10);
        done = done || stop;
        if (!done) eo1 = eo1 + tem5;
        ++ktr;
    } while (ktr <= 10 && !(UNIFORM ? warp_all(done) : done));
    // sin/cos at E + d for the lanes that owe the confirming pass (identity otherwise)
    const double dfin = (done && mag >= kMc[MC_1EM12]) ? tem5 : 0.0;
    const double h = dfin * dfin;
    const double sd = __fma_rn(-0.16666666666666666 * h, dfin, dfin);
    const double cd = __fma_rn(-0.5, h, 1.0);
    const double s_rot = __fma_rn(coseo1, sd, sineo1 * cd);
    const double c_rot = __fma_rn(-sineo1, sd, coseo1 * cd);
    sineo1 = s_rot;
    coseo1 = c_rot;
}

// Kepler's equation for e^2 < 0.64 by Halley's third-order iteration from the same start E = u.
// With f(E) = E - u - g, g = axnl sin E - aynl cos E (= f''), f' = 1 - (axnl cos E + aynl sin E):
// the Newton step dn = -f / f' divided by 1 + dn g / (2 f') (taken only while that factor is
// above 1/2: near the perigee of an eccentric orbit it can change sign, and then the plain
// Newton step is used), clamped like the reference's. Once a step is below 1e-3 it is not applied
// to the iterate: (sin, cos) are rotated by it (Taylor sums, no range reduction), ONE Newton
// correction is formed from the rotated values (|correction| < 1e-8) and applied to first order.
// Passes with a general sin/cos: 2.3-2.7 on average at e = 0.7 (at most 3 for e < 0.8) where the
// reference's iteration takes 5 (measured over 4e5 (u, e) pairs incl. the perigee region); the
// result is within 5e-15 of the root, the reference's own within 1e-12 of it (it stops after a
// correction below 1e-12 that it does not apply). `active` lanes only; a lane stops moving its
// iterate when it is done and the warp leaves on a vote, so a result never depends on the other
// lanes. Returns false for a lane that did not get there in 8 passes (NaN, garbage): those take
// the reference's loop.
template <bool UNIFORM>
__device__ __forceinline__ bool kepler_halley(bool active, double u, double axnl, double aynl,
                                              double &sineo1, double &coseo1) {
    double eo1 = u, se = 0.0, ce = 1.0, d = 0.0;
    bool done = !active, found = false;
    int pass = 0;
#pragma unroll 1
    do {
        sincos_cw(eo1, se, ce);
        const double g = msb2(axnl, se, aynl, ce);
        const double fp = 1.0 - mad2(axnl, ce, aynl, se);
        const double f = (eo1 - u) - g;
        double r;  // 1 / f' to ~2^-40 (a Newton-type step only needs a few digits of it)
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(fp));
        r = __fma_rn(r, __fma_rn(-fp, r, kMc[MC_ONE]), r);
        const double dn = -f * r;
        const double q = __fma_rn(0.5 * dn * g, r, 1.0);
        d = (q > 0.5) ? div_coarse(dn, q) : dn;
        d = (d > kMc[MC_0P95]) ? kMc[MC_0P95] : d;
        d = (d < -kMc[MC_0P95]) ? -kMc[MC_0P95] : d;
        const bool stop = fabs(d) < 0.0009765625;
        found = found || (stop && !done);
        done = done || stop;
        if (!done) eo1 = eo1 + d;
        ++pass;
    } while (pass < 8 && !(UNIFORM ? warp_all(done) : done));
    // (sin, cos) at E + d, then one Newton correction from there
    double sd, cd;
    sincos_taylor<2, 3>(d, sd, cd);
    const double s1 = mad2(ce, sd, se, cd);
    const double c1 = msb2(ce, cd, se, sd);
    const double e1 = eo1 + d;
    const double f1 = (e1 - u) - msb2(axnl, s1, aynl, c1);
    const double fp1 = 1.0 - mad2(axnl, c1, aynl, s1);
    const double d1 = -div_coarse(f1, fp1);
    sineo1 = __fma_rn(c1, d1, s1);
    coseo1 = __fma_rn(-s1, d1, c1);
    return found;
}

// Eccentric orbits (e^2 >= 4e-4): Halley's iteration where e^2 < 0.64, the reference's own
// loop -- whose outcome after its 10 passes is the answer, converged or not -- for the rest and
// for lanes the fast iteration did not settle. Per-lane choices on warp votes (kernel 2).
template <bool UNIFORM>
__device__ __forceinline__ void kepler_eccentric(double u, double axnl, double aynl, double el2,
                                                 double &sineo1, double &coseo1) {
    if (UNIFORM && PB_KEPLER_HALLEY) {
        const bool fast = el2 < 0.64;
        bool ok = false;
        if (warp_ballot(fast) != 0u) {
            ok = kepler_halley<true>(fast, u, axnl, aynl, sineo1, coseo1);
        }
        if (!warp_all(ok)) {
            double sn, cn;
            kepler_newton<true>(u, axnl, aynl, sn, cn);
            sineo1 = ok ? sineo1 : sn;
            coseo1 = ok ? coseo1 : cn;
        }
    } else {
        kepler_newton<UNIFORM>(u, axnl, aynl, sineo1, coseo1);
    }
}

// Kepler's equation for a nearly circular orbit, e^2 = axnl^2 + aynl^2 < 4e-4. With E = u + d
// the equation reads  d = g0 cos d + g1 sin d,  g0 = axnl sin u - aynl cos u,
// g1 = axnl cos u + aynl sin u  (|g0|, |g1| <= e, so |d| <= e / (1 - e)). Its solution to
// fourth order in e is  d4 = g0 ((1 + g1 + g1^2 + g1^3) - g0^2 (1/2 + 5/3 g1))  (|d - d4| <
// 2 e^5 = 7e-9); one Newton step on F(d) = d - g0 cos d - g1 sin d from there, with short Taylor
// sums for sin/cos of the small d (no range reduction, no quadrant logic) and 1 / F' = 1 / (1 -
// w) as 1 + w + w^2 (|w| <= e: the step is below 7e-9, so the w^3 left out moves it by < 6e-14),
// lands within 4e-15 of the root. (sin E, cos E) is (sin u, cos u) rotated by d: ONE general
// sin/cos, where the Newton iteration from E = u needs two or three. The reference stops once
// a correction is below 1e-12 without applying it, so its own E sits within 1e-12 of the same
// root (7e-9 km at LEO): both agree to that, far inside the parity bar.
// (su, cu) = (sin u, cos u) come from the caller: the general routine on u, or rotations of
// sines and cosines the evaluation already holds (sgp4_eval, near-Earth full-drag class).
__device__ __forceinline__ void kepler_small_sc(double su, double cu, double axnl, double aynl,
                                                double &sineo1, double &coseo1) {
    const double g0 = msb2(axnl, su, aynl, cu);
    const double g1 = mad2(axnl, cu, aynl, su);
    const double a = __dadd_rn(__fma_rn(__fma_rn(g1, g1, g1), g1, g1), 1.0);  // 1 + g1 + g1^2 + g1^3
    const double b = __fma_rn(g1, kMc[MC_5O3], kMc[MC_HALF]);
    const double d = __dmul_rn(g0, __fma_rn(-__dmul_rn(g0, g0), b, a));
    double sd, cd;
    sincos_taylor<2, 3>(d, sd, cd);
    const double f = __fma_rn(-g1, sd, __fma_rn(-g0, cd, d));   // F(d)
    const double w = msb2(g1, cd, g0, sd);                      // F'(d) = 1 - w
    const double step = __dmul_rn(-f, __dadd_rn(__fma_rn(w, w, w), 1.0));
    const double s2 = __fma_rn(cd, step, sd);   // sin / cos of d + step, first order:
    const double c2 = __fma_rn(-sd, step, cd);  // step^2 / 2 < 3e-17
    sineo1 = mad2(su, c2, cu, s2);
    coseo1 = msb2(cu, c2, su, s2);
}
__device__ __forceinline__ void kepler_small(double u, double axnl, double aynl, double &sineo1,
                                             double &coseo1) {
    double su, cu;
    sincos_cw(u, su, cu);
    kepler_small_sc(su, cu, axnl, aynl, sineo1, coseo1);
}

// sqrt(1 - x) for 0 <= x < 4e-4 (the nearly circular orbits of kepler_small): binomial series
// through x^4 (next term 7 x^5 / 256 < 3e-19). Every coefficient is an FP64 immediate: four FMAs
// instead of a reciprocal square root seed with two Newton steps, the square-root step and the
// x == 0 select.
__device__ __forceinline__ double sqrt_1m_small(double x) {
    double p = __fma_rn(-0.0390625, x, -0.0625);
    p = __fma_rn(p, x, -0.125);
    p = __fma_rn(p, x, -0.5);
    return __fma_rn(p, x, 1.0);
}

// Returns the raw sgp4 error code (0, 1, 2, 3, 4, 6). out[0..2] = r [km], out[3..5] = v
// [km/s] are written for codes 0 and 6 only, exactly like the reference.
//
// UNIFORM = true (kernel 2): every lane of the warp is active and evaluates the same
// satellite. The function then has NO early returns -- the first failure is latched in `code`
// and the arithmetic simply carries on with garbage that is never stored -- and Kepler's loop
// is driven by a warp vote, so control flow stays warp-uniform from entry to exit. That is
// what lets ptxas keep addresses and FP64 constants on the uniform datapath (LDCU into
// uniform registers, UR-addressed shared loads): the first version's `return`s made everything
// after them "divergent" and cost ~75 LDC + ~50 R2UR + ~100 moves per evaluation (ncu source
// page, profiles/). UNIFORM = false (kernel 1's epoch probe, one thread per satellite) keeps
// plain early returns.
template <int CLS, bool WANT_SIDE, bool UNIFORM, class K>
__device__ __forceinline__ int sgp4_eval(const K &k, const GravConsts &g, bool opsmode_a,
                                         ClassFlags rt, double t, double out[6],
                                         Sgp4Side *side, const ResNodes *nodes = nullptr,
                                         const ThirdBodyPhase *phase = nullptr) {
    int code = 0;
#define PB_FAIL_IF(cond, c)              \
    do {                                 \
        if (UNIFORM) {                   \
            if (code == 0 && (cond)) code = (c); \
        } else if (cond) {               \
            return (c);                  \
        }                                \
    } while (0)
    using T = ClassTraits<CLS>;
    const bool deep = T::kStatic ? T::kDeep : rt.deep;
    const bool simp = T::kStatic ? T::kSimp : rt.isimp;
    const int irez = T::kStatic ? T::kIrez : rt.irez;
    // Near-Earth full-drag orbits (kernel 2, compile-time class): sin/cos of the argument of
    // latitude come from rotations of sines and cosines this evaluation already holds (see the
    // Kepler section); rot_* carry sin/cos(xmdf) and the small angles that were added to it.
    constexpr bool kRotU =
        UNIFORM && (PB_KEPLER_ROT != 0) && T::kStatic && !T::kDeep && !T::kSimp;
    double rot_sm = 0.0, rot_cm = 1.0, rot_eps = 0.0;
    if (WANT_SIDE) {
        side->stage = 0;
        side->resonant = false;
    }

    // ---- secular gravity and atmospheric drag, src/sgp4.cpp:1805-1835 ----
    const double xmdf = k(PF_MO) + k(PF_MDOT) * t;
    const double argpdf = k(PF_ARGPO) + k(PF_ARGPDOT) * t;
    const double nodedf = k(PF_NODEO) + k(PF_NODEDOT) * t;
    double argpm = argpdf;
    double mm = xmdf;
    const double t2 = t * t;
    double nodem = nodedf + k(PF_NODECF) * t2;
    double tempa = 1.0 - k(PF_CC1) * t;
    double tempe = k(PF_BC4) * t;
    double templ = k(PF_T2COF) * t2;

    if (!simp) {
        const double delomg = k(PF_OMGCOF) * t;
        // cos(xmdf) and sin(mm) with mm = xmdf + temp, |temp| small (drag corrections): one
        // sincos of xmdf, then the angle-addition formula with a short series for temp. On
        // B200 an FP64 instruction blocks its scheduler's issue port for both pipe cycles
        // (scripts/ubench/fp64_mix.cu), so a select-heavy "one polynomial for sin or cos"
        // costs more than it saves.
        double sinm, cosm;
        sincos_cw(xmdf, sinm, cosm);
        const double delmtemp = 1.0 + k(PF_ETA) * cosm;
        const double delm = k(PF_XMCOF) * (delmtemp * delmtemp * delmtemp - k(PF_DELMO));
        const double temp = delomg + delm;
        mm = xmdf + temp;
        argpm = argpdf - temp;
        if (kRotU) {
            rot_sm = sinm;
            rot_cm = cosm;
            rot_eps = temp;
        }
        const double t3 = t2 * t;
        const double t4 = t3 * t;
        tempa = tempa - k(PF_D2) * t2 - k(PF_D3) * t3 - k(PF_D4) * t4;
        double sind, cosd;
        sincos_small<6>(temp, sind, cosd);
        const double sin_mm = mad2(sinm, cosd, cosm, sind);
        tempe = tempe + k(PF_BC5) * (sin_mm - k(PF_SINMAO));
        templ = templ + k(PF_T3COF) * t3 + t4 * (k(PF_T4COF) + t * k(PF_T5COF));
    }

    const double no = k(PF_NO);
    double nm = no;
    double em = k(PF_ECCO);
    double inclm = k(PF_INCLO);

    // ---- dspace, src/sgp4.cpp:1018-1166 ----
    if (deep) {
        const double theta = fmod_2pi(k(PF_GSTO) + t * kMc[MC_RPTIM]);  // rptim, :1045
        em = em + k(PF_DEDT) * t;
        inclm = inclm + k(PF_DIDT) * t;
        argpm = argpm + k(PF_DOMDT) * t;
        nodem = nodem + k(PF_DNODT) * t;
        mm = mm + k(PF_DMDT) * t;
        if (irez != 0) {
            // Stateless restart from epoch (== the reference's `atime = 0` branch, :1079-1084).
            // n whole steps, then the partial step ft. With a node table the n steps are one
            // 16 B load; otherwise (kernel 1's epoch probe, or a time outside the table) they
            // are integrated here with the very same operations.
            const int n = resonance_step_count(t);
            const double delt = (t > 0.0) ? 720.0 : -720.0;
            const int kk = (t > 0.0) ? n : -n;
            double atime, xni, xli;
            double xndt, xldot, xnddt;
            const int at = (nodes != nullptr) ? kk - nodes->k0 : -1;
            if (nodes != nullptr && static_cast<unsigned>(at) < static_cast<unsigned>(nodes->count)) {
                // the step kk itself is tabulated (every time of a grid spanning up to 32 days):
                // the pre-pass reached it walking from epoch in the same direction and evaluated
                // resonance_rates there with these very inputs -- same operations, same bits
                const double4 nd = *reinterpret_cast<const double4 *>(nodes->table + (nodes->first + at));
                xli = nd.x;
                xni = nd.y;
                xndt = nd.z;
                xnddt = nd.w;
                xldot = __dadd_rn(xni, k(PF_XFACT));
                atime = __dmul_rn(720.0, static_cast<double>(kk));
            } else {
                // Start from the tabulated node nearest to step kk on the way from epoch (same
                // direction, not beyond kk) or from epoch itself, and walk whatever is left
                // in-line (kernel 1's epoch probe, times outside the table).
                int kj = 0;  // step index the walk starts from (0 = epoch)
                bool from_table = false;
                if (nodes != nullptr && nodes->count > 0) {
                    const int k1 = nodes->k0 + nodes->count - 1;  // last tabulated index
                    if (kk >= 0) {
                        const int cand = kk < k1 ? kk : k1;
                        if (cand >= nodes->k0 && cand >= 0) {
                            kj = cand;
                            from_table = true;
                        }
                    } else {
                        const int cand = kk > nodes->k0 ? kk : nodes->k0;
                        if (cand <= k1 && cand <= 0) {
                            kj = cand;
                            from_table = true;
                        }
                    }
                }
                if (from_table) {
                    const ResNode *nd = nodes->table + (nodes->first + (kj - nodes->k0));
                    xli = nd->xli;
                    xni = nd->xni;
                    atime = __dmul_rn(720.0, static_cast<double>(kj));
                } else {
                    atime = 0.0;
                    xni = no;
                    xli = k(PF_XLAMO);
                }
                const int left = (kk >= 0) ? kk - kj : kj - kk;
                for (int i = 0; i < left; ++i) {
                    resonance_rates(k, irez, xli, xni, atime, xndt, xldot, xnddt);
                    resonance_advance(xndt, xldot, xnddt, delt, xli, xni, atime);
                }
                resonance_rates(k, irez, xli, xni, atime, xndt, xldot, xnddt);
            }
            if (WANT_SIDE) {
                side->atime = atime;
                side->xli = xli;
                side->xni = xni;
                side->resonant = true;
            }
            const double ft = t - atime;
            nm = xni + xndt * ft + xnddt * ft * ft * 0.5;
            const double xl = xli + xldot * ft + xndt * ft * ft * 0.5;
            if (irez != 1) {
                mm = xl - 2.0 * nodem + 2.0 * theta;
            } else {
                mm = xl - nodem - argpm + theta;
            }
            const double dndt = nm - no;
            nm = no + dndt;
        }
    }

    PB_FAIL_IF(nm <= 0.0, 2);  // :1860

    double aobase;
    if (deep && irez != 0) {
        // pow(xke / nm, 2/3) with nm = no (1 + x): AOBASE (1 + x)^(-2/3); the resonance moves
        // the mean motion by 1e-5 .. 1e-3 of itself, so a six-term binomial series replaces
        // cbrt + division; anything larger takes the general route.
        const double x = (nm - no) * rcp_fast(no);
        aobase = k(PF_AOBASE) * pow_m2o3_1p(x);
        if (!(fabs(x) < 0.00390625)) {
            aobase = pow_2o3(div_fast(g.xke, nm));
        }
    } else {
        aobase = k(PF_AOBASE);  // nm == no_unkozai: pow(xke/no, 2/3) is a constant
    }
    const double am = aobase * tempa * tempa;
    const double rsqrt_am = rsqrt_fast(am);
    const double sqrt_am = am * rsqrt_am;
    nm = g.xke * (rsqrt_am * rsqrt_am * rsqrt_am);  // xke / pow(am, 1.5)
    em = em - tempe;
    PB_FAIL_IF(em >= 1.0 || em < -kMc[MC_1EM3], 1);  // :1873
    if (em < kMc[MC_1EM6]) em = kMc[MC_1EM6];
    if (kRotU) {
        const double nt = __dmul_rn(no, templ);
        rot_eps = rot_eps + nt;
        mm = mm + nt;
    } else {
        mm = mm + no * templ;
    }
    const double mm_raw = mm;  // xmdf + drag corrections, unreduced
    double xlm = mm + argpm + nodem;  // (only the WANT_SIDE path uses it)

    if (WANT_SIDE) {
        // the mean elements are handed back to the record (Om, om, mm): the representatives
        // matter, keep the reference's sequence with C fmod semantics (:1884-1887)
        nodem = fmod_2pi(nodem);
        argpm = fmod_2pi(argpm);
        xlm = fmod_2pi(xlm);
        mm = fmod_2pi(xlm - argpm - nodem);
    } else if (!kRotU) {
        // The reference reduces xlm and then mm = xlm - argpm - nodem: modulo 2 pi that IS mm,
        // so ONE exact reduction of mm replaces two reductions and four sums at the magnitude of
        // the unreduced angles (whose rounding, ~1e-13 rad, the reference's result carries).
        // argpm (~0.5 rad a week) and nodem (~0.1 rad a week) stay unreduced: they enter sums
        // that are reduced again and sin/cos, which take any angle. The one place where the
        // node's representative reaches a result, dpper's Lyddane branch (it multiplies nodep
        // by cos(i), :341), reduces it there with C fmod semantics.
        mm = wrap_pi(mm);
    }  // kRotU: the mean anomaly only enters u, which is formed from mm_raw (Kepler section)

    if (WANT_SIDE && code == 0) {
        side->am = am;
        side->em = em;
        side->im = inclm;
        side->Om = nodem;
        side->om = argpm;
        side->mm = mm;
        side->nm = nm;
        side->stage = 1;
    }

    // ---- lunar-solar periodics, src/sgp4.cpp:1908-1958 with dpper :235-368 ----
    double ep = em, xincp = inclm, argpp = argpm, nodep = nodem, mp = mm;
    double sinip, cosip, aycof, xlcof;
    if (deep) {
        // zns, zes, znl, zel of :246-249 (kMc[MC_ZNS], kMc[MC_2ZES] = 2 zes, ...)
        // zf = zm + 2 zes sin(zm): sin/cos(zf) by rotating (sin zm, cos zm) through the small
        // equation-of-centre angle, |2 zes sin zm| <= 0.0335 (Sun) / 0.1098 (Moon), with Taylor
        // sums sized for those bounds (truncation < 5e-18 / 7e-18): no second range reduction.
        double zm, sinzf, coszf, f2, f3, s0, c0, sinzm, coszm, sd, cd;
        if (phase != nullptr) {  // kernel 2: rotated from the CTA's per-satellite / per-time parts
            sinzm = phase->ss;
            coszm = phase->cs;
        } else {
            zm = k(PF_ZMOS) + kMc[MC_ZNS] * t;
            sincos_cw(zm, sinzm, coszm);
        }
        sincos_taylor<3, 4>(kMc[MC_2ZES] * sinzm, sd, cd);
        rotate(sinzm, coszm, sd, cd, sinzf, coszf);
        f2 = 0.5 * sinzf * sinzf - 0.25;
        f3 = -0.5 * sinzf * coszf;
        const double ses = mad2(k(PF_SE2), f2, k(PF_SE3), f3);
        const double sis = mad2(k(PF_SI2), f2, k(PF_SI3), f3);
        const double sls = __fma_rn(k(PF_SL4), sinzf, mad2(k(PF_SL2), f2, k(PF_SL3), f3));
        const double sghs = __fma_rn(k(PF_SGH4), sinzf, mad2(k(PF_SGH2), f2, k(PF_SGH3), f3));
        const double shs = mad2(k(PF_SH2), f2, k(PF_SH3), f3);
        if (phase != nullptr) {
            sinzm = phase->sm;
            coszm = phase->cm;
        } else {
            zm = k(PF_ZMOL) + kMc[MC_ZNL] * t;
            sincos_cw(zm, sinzm, coszm);
        }
        sincos_taylor<4, 5>(kMc[MC_2ZEL] * sinzm, sd, cd);
        rotate(sinzm, coszm, sd, cd, sinzf, coszf);
        f2 = 0.5 * sinzf * sinzf - 0.25;
        f3 = -0.5 * sinzf * coszf;
        const double sel = mad2(k(PF_EE2), f2, k(PF_E3), f3);
        const double sil = mad2(k(PF_XI2), f2, k(PF_XI3), f3);
        const double sll = __fma_rn(k(PF_XL4), sinzf, mad2(k(PF_XL2), f2, k(PF_XL3), f3));
        const double sghl = __fma_rn(k(PF_XGH4), sinzf, mad2(k(PF_XGH2), f2, k(PF_XGH3), f3));
        const double shll = mad2(k(PF_XH2), f2, k(PF_XH3), f3);
        // peo, pinco, plo, pgho, pho are zero after dscom (:494-498)
        const double pe = ses + sel;
        const double pinc = sis + sil;
        const double pl = sls + sll;
        double pgh = sghs + sghl;
        double ph = shs + shll;
        xincp = xincp + pinc;
        ep = ep + pe;
        // xincp = inclo + (didt t + pinc): rotate the per-satellite (sin, cos)(inclo) through
        // the small secular + periodic change (1e-3 rad over a week; general routine beyond 2^-6)
        sincos_small<6>(xincp - k(PF_INCLO), sd, cd);
        rotate(k(PF_SINIO), k(PF_COSIO), sd, cd, s0, c0);
        if (xincp >= kMc[MC_0P2]) {  // :317
            ph = div_fast(ph, s0);
            pgh = pgh - c0 * ph;
            argpp = argpp + pgh;
            nodep = nodep + ph;
            mp = mp + pl;
        } else {  // Lyddane modification, :325-364
            // The reference forms alfdp = A sin(node) + ph cos(node), betdp = A cos(node) -
            // ph sin(node) with A = sin(i) + pinc cos(i) and takes atan2(alfdp, betdp), then
            // moves the result by +-2pi onto the branch within pi of the old node: that IS
            // node + atan2(ph, A). No sin/cos of the node, and for |ph| < A / 4 (every
            // satellite but those within ~1e-3 rad of the equator) a 1-ulp polynomial arctangent
            // of the small ratio instead of a general atan2.
            const double A = s0 + pinc * c0;
            const double tq = div_fast(ph, A);
            double dnode = atan_small(tq);
            if (!(fabs(ph) < 0.25 * A)) {  // per lane: this branch is not warp-uniform anyway
                dnode = atan2(ph, A);
            }
            nodep = fmod_2pi(nodep);
            if (nodep < 0.0 && opsmode_a) nodep = nodep + kMc[MC_TWOPI];
            double xls = mp + argpp + c0 * nodep;
            const double dls = pl + pgh - pinc * nodep * s0;
            xls = xls + dls;
            nodep = nodep + dnode;
            mp = mp + pl;
            argpp = xls - mp - c0 * nodep;
        }
        // sin/cos of the perturbed inclination are already known from dpper (s0, c0); the
        // reference recomputes them after the sign fix (:1950-1951): sin(-x) = -sin(x).
        sinip = s0;
        cosip = c0;
        if (xincp < 0.0) {  // :1932
            xincp = -xincp;
            nodep = nodep + kMc[MC_PI];
            argpp = argpp - kMc[MC_PI];
            sinip = -s0;
        }
        PB_FAIL_IF(ep < 0.0 || ep > 1.0, 3);  // :1938
        aycof = -0.5 * g.j3oj2 * sinip;
        {   // :1954-1957: the denominator is 1 + cos(i), or 1.5e-12 within that of a retrograde
            // equatorial orbit -- a select, not a branch
            const double den = (fabs(cosip + 1.0) > kMc[MC_1P5EM12]) ? 1.0 + cosip : kMc[MC_1P5EM12];
            xlcof = div_fast(-0.25 * g.j3oj2 * sinip * (3.0 + 5.0 * cosip), den);
        }
        if (WANT_SIDE && code == 0) {
            side->aycof = aycof;
            side->xlcof = xlcof;
            side->stage = 2;
        }
    } else {
        sinip = k(PF_SINIO);
        cosip = k(PF_COSIO);
        aycof = k(PF_AYCOF);
        xlcof = k(PF_XLCOF);
    }

    // ---- long-period periodics and Kepler's equation, src/sgp4.cpp:1959-1983 ----
    double sinargp, cosargp;
    sincos_cw(argpp, sinargp, cosargp);
    const double axnl = ep * cosargp;
    double temp = rcp_fast(am * (1.0 - ep * ep));
    const double aynl = mad2(ep, sinargp, temp, aycof);
    const double el2 = mad2(axnl, axnl, aynl, aynl);
    // Kepler's equation  E = u - aynl cos E + axnl sin E  (src/sgp4.cpp:1965-1983). Nearly
    // circular orbits (el2 < 4e-4, i.e. e < 0.02: most of any catalogue) are solved for
    // delta = E - u in closed form, one general sin/cos instead of two or three (kepler_small);
    // everything else runs the reference's Newton iteration (kepler_newton). The choice is made
    // per lane from the lane's own el2 -- a result never depends on the other lanes of the warp
    // -- while the branches are taken on warp votes, so control flow stays warp-uniform.
    // betal = sqrt(1 - el2) (:1998) follows the same split: a four-term series where el2 < 4e-4.
    double sineo1 = 0.0, coseo1 = 1.0, betal = 1.0;
    if (kRotU) {
        // u = xl - nodep = mp + argpp + dxl (mod 2 pi) with mp = xmdf + (drag corrections) and
        // sin/cos of xmdf and of argpp both at hand: (sin u, cos u) is (sin, cos)(xmdf) rotated
        // by argpp and then by the small angle eps = drag corrections + dxl (a few 1e-3 rad after
        // a day, ~1e-2 after a week: Taylor sums, truncation < 3e-19 below 2^-4). No range
        // reduction of xlm, mm and u (the reference's three fmod calls, :1884-1887, 1963) and no
        // general sin/cos of u: 19 FP64 operations instead of 36, no quadrant logic. A lane whose
        // eps is larger (strong drag, weeks from epoch) takes the general route with the lanes
        // that are not nearly circular; u itself is only formed there, as Newton's starting value.
        const double dxl = temp * xlcof * axnl;
        const double eps = rot_eps + dxl;
        const bool small = (el2 < kMc[MC_KEPLER_SMALL]) && (fabs(eps) < 0.0625);
        const unsigned small_lanes = warp_ballot(small);
        if (small_lanes != 0u) {
            double s1, c1, se, ce, su, cu;
            rotate(rot_sm, rot_cm, sinargp, cosargp, s1, c1);
            sincos_taylor<4, 4>(eps, se, ce);
            rotate(s1, c1, se, ce, su, cu);
            kepler_small_sc(su, cu, axnl, aynl, sineo1, coseo1);
            betal = sqrt_1m_small(el2);
        }
        if (small_lanes != 0xffffffffu) {
            const double u = wrap_pi((mm_raw + argpp) + dxl);
            double sn, cn;
            kepler_eccentric<true>(u, axnl, aynl, el2, sn, cn);
            const double bn = sqrt_fast(1.0 - el2);
            sineo1 = small ? sineo1 : sn;
            coseo1 = small ? coseo1 : cn;
            betal = small ? betal : bn;
        }
    } else {
        const double xl = mp + argpp + nodep + temp * xlcof * axnl;
        const double u = wrap_pi(xl - nodep);  // fmod(xl - nodep, twopi); only u mod 2pi matters
        if (UNIFORM && PB_KEPLER_SMALL) {
            const bool small = el2 < kMc[MC_KEPLER_SMALL];
            // ONE vote: both branch decisions below derive from the same warp-uniform mask
            const unsigned small_lanes = warp_ballot(small);
            if (small_lanes == 0xffffffffu) {
                kepler_small(u, axnl, aynl, sineo1, coseo1);
                betal = sqrt_1m_small(el2);
            } else {
                kepler_eccentric<true>(u, axnl, aynl, el2, sineo1, coseo1);
                betal = sqrt_fast(1.0 - el2);
                if (small_lanes != 0u) {
                    double sf, cf;
                    kepler_small(u, axnl, aynl, sf, cf);
                    sineo1 = small ? sf : sineo1;
                    coseo1 = small ? cf : coseo1;
                    betal = small ? sqrt_1m_small(el2) : betal;
                }
            }
        } else {
            kepler_newton<UNIFORM>(u, axnl, aynl, sineo1, coseo1);
            betal = sqrt_fast(1.0 - el2);
        }
    }

    // ---- short-period preliminary quantities, src/sgp4.cpp:1986-2011 ----
    const double ecose = mad2(axnl, coseo1, aynl, sineo1);
    const double esine = msb2(axnl, sineo1, aynl, coseo1);
    const double pl = am * (1.0 - el2);
    PB_FAIL_IF(pl < 0.0, 4);  // :1990

    const double rl = am * (1.0 - ecose);
    const double inv_rl = rcp_fast(rl);
    const double rdotl = sqrt_am * esine * inv_rl;
    const double rvdotl = (sqrt_am * betal) * inv_rl;  // sqrt(pl) = sqrt(am) sqrt(1 - el2)
    temp = esine * rcp_fast(1.0 + betal);
    const double aorl = am * inv_rl;
    const double sinu = aorl * (sineo1 - aynl - axnl * temp);
    const double cosu = aorl * (coseo1 - axnl + aynl * temp);
    const double sin2u = (cosu + cosu) * sinu;
    const double cos2u = 1.0 - 2.0 * sinu * sinu;
    temp = rcp_fast(pl);
    const double temp1 = 0.5 * g.j2 * temp;
    const double temp2 = temp1 * temp;

    double con41, x1mth2, x7thm1;
    if (deep) {  // :2014-2020
        const double cosisq = cosip * cosip;
        con41 = 3.0 * cosisq - 1.0;
        x1mth2 = 1.0 - cosisq;
        x7thm1 = 7.0 * cosisq - 1.0;
        if (WANT_SIDE && code == 0) {
            side->con41 = con41;
            side->x1mth2 = x1mth2;
            side->x7thm1 = x7thm1;
            side->stage = 3;
        }
    } else {
        con41 = k(PF_CON41);
        x1mth2 = k(PF_X1MTH2);
        x7thm1 = k(PF_X7THM1);
    }
    const double mrt = mad2(rl, 1.0 - 1.5 * temp2 * betal * con41, 0.5 * temp1 * x1mth2, cos2u);
    // The reference forms su = atan2(sinu, cosu) - dsu and then takes sin(su), cos(su)
    // (:2006, 2023, 2031-2032). Only those two are used, so rotate the unit vector
    // (sinu, cosu)/hypot by the small angle dsu instead: no atan2, no second range reduction.
    const double dsu = 0.25 * temp2 * x7thm1 * sin2u;
    const double xnode = nodep + 1.5 * temp2 * cosip * sin2u;
    const double dxi = 1.5 * temp2 * cosip * sinip * cos2u;  // xinc = xincp + dxi
    const double nmt = nm * temp1 * g.inv_xke;
    const double mvt = rdotl - nmt * x1mth2 * sin2u;
    const double rvdot = rvdotl + nmt * mad2(x1mth2, cos2u, 1.5, con41);

    // ---- orientation vectors, position and velocity, src/sgp4.cpp:2031-2052 ----
    double sinsu, cossu, snod, cnod, sini, cosi;
    {
        // (sinu, cosu) is a unit vector by construction: with r = a (1 - e cos E) and
        // beta = sqrt(1 - e^2), sinu^2 + cosu^2 == 1 is an algebraic identity in E (any E,
        // converged or not), so atan2 followed by sin/cos would only remove rounding noise of
        // a few 1e-16 (times a/r near the perigee of a very eccentric orbit) from the norm.
        // xinc = xincp + dxi likewise, with sin/cos(xincp) already at hand. Both corrections are
        // ~1e-3 rad for every physical orbit: two-coefficient Taylor sums, and ONE fall-back
        // region for the garbage states where either is not small (decided per lane).
        double sd, cd, sdx, cdx;
        sincos_series<9>(dsu, sd, cd);
        sincos_series<9>(dxi, sdx, cdx);
        if (!(fabs(dsu) < 0.001953125 && fabs(dxi) < 0.001953125)) {
            sincos_cw(dsu, sd, cd);
            sincos_cw(dxi, sdx, cdx);
        }
        sinsu = msb2(sinu, cd, cosu, sd);
        cossu = mad2(cosu, cd, sinu, sd);
        sini = mad2(sinip, cdx, cosip, sdx);
        cosi = msb2(cosip, cdx, sinip, sdx);
    }
    sincos_cw(xnode, snod, cnod);
    const double xmx = -snod * cosi;
    const double xmy = cnod * cosi;
    const double ux = mad2(xmx, sinsu, cnod, cossu);
    const double uy = mad2(xmy, sinsu, snod, cossu);
    const double uz = sini * sinsu;
    const double vx = msb2(xmx, cossu, cnod, sinsu);
    const double vy = msb2(xmy, cossu, snod, sinsu);
    const double vz = sini * cossu;

    const double mr = mrt * g.radiusearthkm;
    const double mv = mvt * g.vkmpersec, rv = rvdot * g.vkmpersec;
    out[0] = mr * ux;
    out[1] = mr * uy;
    out[2] = mr * uz;
    out[3] = mad2(mv, ux, rv, vx);
    out[4] = mad2(mv, uy, rv, vy);
    out[5] = mad2(mv, uz, rv, vz);

    PB_FAIL_IF(mrt < 1.0, 6);  // :2056
#undef PB_FAIL_IF
    return code;
}

}  // namespace pb

#endif
