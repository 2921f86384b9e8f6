// Kernel 1 core: per-satellite SGP4 initialisation on the device. Counterpart of
//   Satellite(const TwoLineElement&, GravModel)  src/perturb.cpp:135-186  (units, epoch)
//   sgp4::sgp4init                               src/sgp4.cpp:1383-1680
//     initl  :1221-1299   dscom :438-636   dpper(init='y') :235-368   dsinit :719-943
//   gstime_SGP4 :2506-2525   days2mdhms_SGP4 :3161-3192   jday_SGP4 :3100-3121
// The translation unit that includes this header is compiled with -fmad=false: every
// multiply/add below then rounds exactly like the reference's scalar code and only the libm
// calls (sin, cos, pow, fmod, atan2 from libdevice vs glibc) can differ, by an ulp or two.
// Throughput is irrelevant here (about 1 us per satellite once), fidelity is not: the
// near/deep, isimp, irez and perigee thresholds are evaluated on these values.
#ifndef PERTURB_B200_SGP4_INIT_CUH
#define PERTURB_B200_SGP4_INIT_CUH

#include "sgp4_eval.cuh"
#include "sgp4_fields.h"

namespace pb {

struct GravModelConsts {  // getgravconst(), src/sgp4.cpp:2101-2157 (filled on the host)
    double tumin, mus, radiusearthkm, xke, j2, j3, j4, j3oj2;
};

struct TleNumbers {  // perturb::TwoLineElement numerics, include/perturb/tle.hpp:65-91
    double epoch_year, epoch_day_of_year, n_dot, n_ddot, b_star;
    double inclination, raan, eccentricity, arg_of_perigee, mean_anomaly, mean_motion;
};

struct SatRec {
    double f[RF_COUNT];
};

// src/sgp4.cpp:2506-2525
__device__ inline double gstime_dev(double jdut1) {
    const double deg2rad = kPi / 180.0;
    const double tut1 = (jdut1 - 2451545.0) / 36525.0;
    double temp = -6.2e-6 * tut1 * tut1 * tut1 + 0.093104 * tut1 * tut1 +
                  (876600.0 * 3600 + 8640184.812866) * tut1 + 67310.54841;
    temp = fmod(temp * deg2rad / 240.0, kTwoPi);
    if (temp < 0.0) temp += kTwoPi;
    return temp;
}

// days2mdhms_SGP4 + jday_SGP4 as Satellite(TwoLineElement) chains them, perturb.cpp:170-178
__device__ inline void epoch_to_jd(double epoch_year, double epoch_days, double &jd,
                                   double &jd_frac) {
    const int yy = static_cast<int>(epoch_year);
    const int year = yy + ((yy < 57) ? 2000 : 1900);
    int mlen[13] = {0, 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    if ((year % 4) == 0) mlen[2] = 29;
    const int dayofyr = static_cast<int>(floor(epoch_days));
    int mon = 1, acc = 0;
    while ((dayofyr > acc + mlen[mon]) && (mon < 12)) {
        acc += mlen[mon];
        mon++;
    }
    const int day = dayofyr - acc;
    double temp = (epoch_days - dayofyr) * 24.0;
    const int hr = static_cast<int>(floor(temp));
    temp = (temp - hr) * 60.0;
    const int minute = static_cast<int>(floor(temp));
    const double sec = (temp - minute) * 60.0;

    jd = 367.0 * year - floor((7 * (year + floor((mon + 9) / 12.0))) * 0.25) +
         floor(275 * mon / 9.0) + day + 1721013.5;
    jd_frac = (sec + minute * 60.0 + hr * 3600.0) / 86400.0;
    if (fabs(jd_frac) > 1.0) {
        const double dtt = floor(jd_frac);
        jd = jd + dtt;
        jd_frac = jd_frac - dtt;
    }
}

struct BodyTerms {  // one pass of the dscom loop body, src/sgp4.cpp:527-570
    double s1, s2, s3, s4, s5, s6, s7;
    double z1, z2, z3, z11, z12, z13, z21, z22, z23, z31, z32, z33;
};

struct OrbitGeom {
    double sinim, cosim, sinomm, cosomm, snodm, cnodm, em, emsq, betasq, rtemsq, xnoi;
};

__device__ inline BodyTerms third_body_terms(const OrbitGeom &o, double zcosg, double zsing,
                                             double zcosi, double zsini, double zcosh,
                                             double zsinh, double cc) {
    BodyTerms b;
    const double a1 = zcosg * zcosh + zsing * zcosi * zsinh;
    const double a3 = -zsing * zcosh + zcosg * zcosi * zsinh;
    const double a7 = -zcosg * zsinh + zsing * zcosi * zcosh;
    const double a8 = zsing * zsini;
    const double a9 = zsing * zsinh + zcosg * zcosi * zcosh;
    const double a10 = zcosg * zsini;
    const double a2 = o.cosim * a7 + o.sinim * a8;
    const double a4 = o.cosim * a9 + o.sinim * a10;
    const double a5 = -o.sinim * a7 + o.cosim * a8;
    const double a6 = -o.sinim * a9 + o.cosim * a10;
    const double x1 = a1 * o.cosomm + a2 * o.sinomm;
    const double x2 = a3 * o.cosomm + a4 * o.sinomm;
    const double x3 = -a1 * o.sinomm + a2 * o.cosomm;
    const double x4 = -a3 * o.sinomm + a4 * o.cosomm;
    const double x5 = a5 * o.sinomm;
    const double x6 = a6 * o.sinomm;
    const double x7 = a5 * o.cosomm;
    const double x8 = a6 * o.cosomm;
    const double e2 = o.emsq;
    b.z31 = 12.0 * x1 * x1 - 3.0 * x3 * x3;
    b.z32 = 24.0 * x1 * x2 - 6.0 * x3 * x4;
    b.z33 = 12.0 * x2 * x2 - 3.0 * x4 * x4;
    b.z1 = 3.0 * (a1 * a1 + a2 * a2) + b.z31 * e2;
    b.z2 = 6.0 * (a1 * a3 + a2 * a4) + b.z32 * e2;
    b.z3 = 3.0 * (a3 * a3 + a4 * a4) + b.z33 * e2;
    b.z11 = -6.0 * a1 * a5 + e2 * (-24.0 * x1 * x7 - 6.0 * x3 * x5);
    b.z12 = -6.0 * (a1 * a6 + a3 * a5) +
            e2 * (-24.0 * (x2 * x7 + x1 * x8) - 6.0 * (x3 * x6 + x4 * x5));
    b.z13 = -6.0 * a3 * a6 + e2 * (-24.0 * x2 * x8 - 6.0 * x4 * x6);
    b.z21 = 6.0 * a2 * a5 + e2 * (24.0 * x1 * x5 - 6.0 * x3 * x7);
    b.z22 = 6.0 * (a4 * a5 + a2 * a6) +
            e2 * (24.0 * (x2 * x5 + x1 * x6) - 6.0 * (x4 * x7 + x3 * x8));
    b.z23 = 6.0 * a4 * a6 + e2 * (24.0 * x2 * x6 - 6.0 * x4 * x8);
    b.z1 = b.z1 + b.z1 + o.betasq * b.z31;
    b.z2 = b.z2 + b.z2 + o.betasq * b.z32;
    b.z3 = b.z3 + b.z3 + o.betasq * b.z33;
    b.s3 = cc * o.xnoi;
    b.s2 = -0.5 * b.s3 / o.rtemsq;
    b.s4 = b.s3 * o.rtemsq;
    b.s1 = -15.0 * o.em * b.s4;
    b.s5 = x1 * x3 + x2 * x4;
    b.s6 = x2 * x3 + x1 * x4;
    b.s7 = x2 * x4 - x1 * x3;
    return b;
}

#define R(name) rec.f[RF_##name]

// Deep-space set-up: dscom + dsinit (dpper(init='y') only evaluates periodics that are then
// discarded -- with init=='y' nothing is applied, src/sgp4.cpp:296 -- so it has no effect on
// the record and is not reproduced).
__device__ inline void deep_space_setup(SatRec &rec, double epoch, double eccsq, double xpidot) {
    constexpr double zes = 0.01675, zel = 0.05490, zns = 1.19459e-5, znl = 1.5835218e-4;
    constexpr double rptim = 4.37526908801129966e-3;
    OrbitGeom o;
    o.snodm = sin(R(nodeo));
    o.cnodm = cos(R(nodeo));
    o.sinomm = sin(R(argpo));
    o.cosomm = cos(R(argpo));
    o.sinim = sin(R(inclo));
    o.cosim = cos(R(inclo));
    o.em = R(ecco);
    o.emsq = o.em * o.em;
    o.betasq = 1.0 - o.emsq;
    o.rtemsq = sqrt(o.betasq);
    o.xnoi = 1.0 / R(no_unkozai);

    const double tc = 0.0;
    const double day = epoch + 18261.5 + tc / 1440.0;
    const double xnodce = fmod(4.5236020 - 9.2422029e-4 * day, kTwoPi);
    const double stem = sin(xnodce);
    const double ctem = cos(xnodce);
    const double zcosil = 0.91375164 - 0.03568096 * ctem;
    const double zsinil = sqrt(1.0 - zcosil * zcosil);
    const double zsinhl = 0.089683511 * stem / zsinil;
    const double zcoshl = sqrt(1.0 - zsinhl * zsinhl);
    const double gam = 5.8351514 + 0.0019443680 * day;
    double zx = 0.39785416 * stem / zsinil;
    const double zy = zcoshl * ctem + 0.91744867 * zsinhl * stem;
    zx = atan2(zx, zy);
    zx = gam + zx - xnodce;
    const double zcosgl = cos(zx);
    const double zsingl = sin(zx);

    const BodyTerms sun = third_body_terms(o, 0.1945905, -0.98088458, 0.91744867, 0.39785416,
                                           o.cnodm, o.snodm, 2.9864797e-6);
    const BodyTerms moon = third_body_terms(
        o, zcosgl, zsingl, zcosil, zsinil, zcoshl * o.cnodm + zsinhl * o.snodm,
        o.snodm * zcoshl - o.cnodm * zsinhl, 4.7968065e-7);

    R(zmol) = fmod(4.7199672 + 0.22997150 * day - gam, kTwoPi);
    R(zmos) = fmod(6.2565837 + 0.017201977 * day, kTwoPi);

    R(se2) = 2.0 * sun.s1 * sun.s6;
    R(se3) = 2.0 * sun.s1 * sun.s7;
    R(si2) = 2.0 * sun.s2 * sun.z12;
    R(si3) = 2.0 * sun.s2 * (sun.z13 - sun.z11);
    R(sl2) = -2.0 * sun.s3 * sun.z2;
    R(sl3) = -2.0 * sun.s3 * (sun.z3 - sun.z1);
    R(sl4) = -2.0 * sun.s3 * (-21.0 - 9.0 * o.emsq) * zes;
    R(sgh2) = 2.0 * sun.s4 * sun.z32;
    R(sgh3) = 2.0 * sun.s4 * (sun.z33 - sun.z31);
    R(sgh4) = -18.0 * sun.s4 * zes;
    R(sh2) = -2.0 * sun.s2 * sun.z22;
    R(sh3) = -2.0 * sun.s2 * (sun.z23 - sun.z21);

    R(ee2) = 2.0 * moon.s1 * moon.s6;
    R(e3) = 2.0 * moon.s1 * moon.s7;
    R(xi2) = 2.0 * moon.s2 * moon.z12;
    R(xi3) = 2.0 * moon.s2 * (moon.z13 - moon.z11);
    R(xl2) = -2.0 * moon.s3 * moon.z2;
    R(xl3) = -2.0 * moon.s3 * (moon.z3 - moon.z1);
    R(xl4) = -2.0 * moon.s3 * (-21.0 - 9.0 * o.emsq) * zel;
    R(xgh2) = 2.0 * moon.s4 * moon.z32;
    R(xgh3) = 2.0 * moon.s4 * (moon.z33 - moon.z31);
    R(xgh4) = -18.0 * moon.s4 * zel;
    R(xh2) = -2.0 * moon.s2 * moon.z22;
    R(xh3) = -2.0 * moon.s2 * (moon.z23 - moon.z21);

    // ---- dsinit ----
    const double nm = R(no_unkozai);
    const double inclm = R(inclo);
    const double cosim = o.cosim, sinim = o.sinim;
    double em = o.em, emsq = o.emsq;

    int irez = 0;
    if ((nm < 0.0052359877) && (nm > 0.0034906585)) irez = 1;
    if ((nm >= 8.26e-3) && (nm <= 9.24e-3) && (em >= 0.5)) irez = 2;
    R(irez) = irez;

    const double ses = sun.s1 * zns * sun.s5;
    const double sis = sun.s2 * zns * (sun.z11 + sun.z13);
    const double sls = -zns * sun.s3 * (sun.z1 + sun.z3 - 14.0 - 6.0 * emsq);
    const double sghs = sun.s4 * zns * (sun.z31 + sun.z33 - 6.0);
    double shs = -zns * sun.s2 * (sun.z21 + sun.z23);
    const bool polar = (inclm < 5.2359877e-2) || (inclm > kPi - 5.2359877e-2);
    if (polar) shs = 0.0;
    if (sinim != 0.0) shs = shs / sinim;
    const double sgs = sghs - cosim * shs;

    R(dedt) = ses + moon.s1 * znl * moon.s5;
    R(didt) = sis + moon.s2 * znl * (moon.z11 + moon.z13);
    R(dmdt) = sls - znl * moon.s3 * (moon.z1 + moon.z3 - 14.0 - 6.0 * emsq);
    const double sghl = moon.s4 * znl * (moon.z31 + moon.z33 - 6.0);
    double shll = -znl * moon.s2 * (moon.z21 + moon.z23);
    if (polar) shll = 0.0;
    R(domdt) = sgs + sghl;
    R(dnodt) = shs;
    if (sinim != 0.0) {
        R(domdt) = R(domdt) - cosim / sinim * shll;
        R(dnodt) = R(dnodt) + shll / sinim;
    }

    if (irez == 0) return;

    const double theta = fmod(R(gsto) + tc * rptim, kTwoPi);
    const double aonv = pow(nm / R(xke), 2.0 / 3.0);

    if (irez == 2) {
        const double cosisq = cosim * cosim;
        em = R(ecco);
        emsq = eccsq;
        const double eoc = em * emsq;
        const double g201 = -0.306 - (em - 0.64) * 0.440;
        double g211, g310, g322, g410, g422, g520, g533, g521, g532;
        if (em <= 0.65) {
            g211 = 3.616 - 13.2470 * em + 16.2900 * emsq;
            g310 = -19.302 + 117.3900 * em - 228.4190 * emsq + 156.5910 * eoc;
            g322 = -18.9068 + 109.7927 * em - 214.6334 * emsq + 146.5816 * eoc;
            g410 = -41.122 + 242.6940 * em - 471.0940 * emsq + 313.9530 * eoc;
            g422 = -146.407 + 841.8800 * em - 1629.014 * emsq + 1083.4350 * eoc;
            g520 = -532.114 + 3017.977 * em - 5740.032 * emsq + 3708.2760 * eoc;
        } else {
            g211 = -72.099 + 331.819 * em - 508.738 * emsq + 266.724 * eoc;
            g310 = -346.844 + 1582.851 * em - 2415.925 * emsq + 1246.113 * eoc;
            g322 = -342.585 + 1554.908 * em - 2366.899 * emsq + 1215.972 * eoc;
            g410 = -1052.797 + 4758.686 * em - 7193.992 * emsq + 3651.957 * eoc;
            g422 = -3581.690 + 16178.110 * em - 24462.770 * emsq + 12422.520 * eoc;
            if (em > 0.715) {
                g520 = -5149.66 + 29936.92 * em - 54087.36 * emsq + 31324.56 * eoc;
            } else {
                g520 = 1464.74 - 4664.75 * em + 3763.64 * emsq;
            }
        }
        if (em < 0.7) {
            g533 = -919.22770 + 4988.6100 * em - 9064.7700 * emsq + 5542.21 * eoc;
            g521 = -822.71072 + 4568.6173 * em - 8491.4146 * emsq + 5337.524 * eoc;
            g532 = -853.66600 + 4690.2500 * em - 8624.7700 * emsq + 5341.4 * eoc;
        } else {
            g533 = -37995.780 + 161616.52 * em - 229838.20 * emsq + 109377.94 * eoc;
            g521 = -51752.104 + 218913.95 * em - 309468.16 * emsq + 146349.42 * eoc;
            g532 = -40023.880 + 170470.89 * em - 242699.48 * emsq + 115605.82 * eoc;
        }
        const double sini2 = sinim * sinim;
        const double f220 = 0.75 * (1.0 + 2.0 * cosim + cosisq);
        const double f221 = 1.5 * sini2;
        const double f321 = 1.875 * sinim * (1.0 - 2.0 * cosim - 3.0 * cosisq);
        const double f322 = -1.875 * sinim * (1.0 + 2.0 * cosim - 3.0 * cosisq);
        const double f441 = 35.0 * sini2 * f220;
        const double f442 = 39.3750 * sini2 * sini2;
        const double f522 = 9.84375 * sinim *
                            (sini2 * (1.0 - 2.0 * cosim - 5.0 * cosisq) +
                             0.33333333 * (-2.0 + 4.0 * cosim + 6.0 * cosisq));
        const double f523 = sinim * (4.92187512 * sini2 * (-2.0 - 4.0 * cosim + 10.0 * cosisq) +
                                     6.56250012 * (1.0 + 2.0 * cosim - 3.0 * cosisq));
        const double f542 = 29.53125 * sinim *
                            (2.0 - 8.0 * cosim + cosisq * (-12.0 + 8.0 * cosim + 10.0 * cosisq));
        const double f543 = 29.53125 * sinim *
                            (-2.0 - 8.0 * cosim + cosisq * (12.0 + 8.0 * cosim - 10.0 * cosisq));
        const double xno2 = nm * nm;
        const double ainv2 = aonv * aonv;
        double temp1 = 3.0 * xno2 * ainv2;
        double temp = temp1 * 1.7891679e-6;
        R(d2201) = temp * f220 * g201;
        R(d2211) = temp * f221 * g211;
        temp1 = temp1 * aonv;
        temp = temp1 * 3.7393792e-7;
        R(d3210) = temp * f321 * g310;
        R(d3222) = temp * f322 * g322;
        temp1 = temp1 * aonv;
        temp = 2.0 * temp1 * 7.3636953e-9;
        R(d4410) = temp * f441 * g410;
        R(d4422) = temp * f442 * g422;
        temp1 = temp1 * aonv;
        temp = temp1 * 1.1428639e-7;
        R(d5220) = temp * f522 * g520;
        R(d5232) = temp * f523 * g532;
        temp = 2.0 * temp1 * 2.1765803e-9;
        R(d5421) = temp * f542 * g521;
        R(d5433) = temp * f543 * g533;
        R(xlamo) = fmod(R(mo) + R(nodeo) + R(nodeo) - theta - theta, kTwoPi);
        R(xfact) = R(mdot) + R(dmdt) + 2.0 * (R(nodedot) + R(dnodt) - rptim) - R(no_unkozai);
        em = o.em;
        emsq = o.emsq;
    }

    if (irez == 1) {
        const double g200 = 1.0 + emsq * (-2.5 + 0.8125 * emsq);
        const double g310 = 1.0 + 2.0 * emsq;
        const double g300 = 1.0 + emsq * (-6.0 + 6.60937 * emsq);
        const double f220 = 0.75 * (1.0 + cosim) * (1.0 + cosim);
        const double f311 = 0.9375 * sinim * sinim * (1.0 + 3.0 * cosim) - 0.75 * (1.0 + cosim);
        double f330 = 1.0 + cosim;
        f330 = 1.875 * f330 * f330 * f330;
        R(del1) = 3.0 * nm * nm * aonv * aonv;
        R(del2) = 2.0 * R(del1) * f220 * g200 * 1.7891679e-6;
        R(del3) = 3.0 * R(del1) * f330 * g300 * 2.2123015e-7 * aonv;
        R(del1) = R(del1) * f311 * g310 * 2.1460748e-6 * aonv;
        R(xlamo) = fmod(R(mo) + R(nodeo) + R(argpo) - theta, kTwoPi);
        R(xfact) = R(mdot) + xpidot - rptim + R(dmdt) + R(domdt) + R(dnodt) - R(no_unkozai);
    }

    R(xli) = R(xlamo);
    R(xni) = R(no_unkozai);
    R(atime) = 0.0;
}

// sgp4init up to (not including) the trailing sgp4(t=0) call. `rec` must arrive zeroed
// except jdsatepoch / jdsatepochF.
__device__ inline void sgp4init_core(SatRec &rec, const GravModelConsts &gc, double epoch,
                                     double bstar, double ndot, double nddot, double ecco,
                                     double argpo, double inclo, double mo, double no_kozai,
                                     double nodeo) {
    R(tumin) = gc.tumin;
    R(mus) = gc.mus;
    R(radiusearthkm) = gc.radiusearthkm;
    R(xke) = gc.xke;
    R(j2) = gc.j2;
    R(j3) = gc.j3;
    R(j4) = gc.j4;
    R(j3oj2) = gc.j3oj2;
    R(bstar) = bstar;
    R(ndot) = ndot;
    R(nddot) = nddot;
    R(ecco) = ecco;
    R(argpo) = argpo;
    R(inclo) = inclo;
    R(mo) = mo;
    R(no_kozai) = no_kozai;
    R(nodeo) = nodeo;

    const double ss = 78.0 / gc.radiusearthkm + 1.0;
    const double qzms2ttemp = (120.0 - 78.0) / gc.radiusearthkm;
    const double qzms2t = qzms2ttemp * qzms2ttemp * qzms2ttemp * qzms2ttemp;
    const double x2o3 = 2.0 / 3.0;

    // ---- initl ----
    const double eccsq = ecco * ecco;
    const double omeosq = 1.0 - eccsq;
    const double rteosq = sqrt(omeosq);
    const double cosio = cos(inclo);
    const double cosio2 = cosio * cosio;
    const double ak = pow(gc.xke / no_kozai, x2o3);
    const double d1 = 0.75 * gc.j2 * (3.0 * cosio2 - 1.0) / (rteosq * omeosq);
    double del = d1 / (ak * ak);
    const double adel = ak * (1.0 - del * del - del * (1.0 / 3.0 + 134.0 * del * del / 81.0));
    del = d1 / (adel * adel);
    const double no = no_kozai / (1.0 + del);
    R(no_unkozai) = no;
    const double ao = pow(gc.xke / no, x2o3);
    const double sinio = sin(inclo);
    const double po = ao * omeosq;
    const double con42 = 1.0 - 5.0 * cosio2;
    R(con41) = -con42 - cosio2 - cosio2;
    const double posq = po * po;
    const double rp = ao * (1.0 - ecco);
    R(method_d) = 0.0;
    R(gsto) = gstime_dev(epoch + 2433281.5);

    R(a) = pow(no * gc.tumin, (-2.0 / 3.0));
    R(alta) = R(a) * (1.0 + ecco) - 1.0;
    R(altp) = R(a) * (1.0 - ecco) - 1.0;

    if ((omeosq >= 0.0) || (no >= 0.0)) {
        R(isimp) = 0.0;
        if (rp < (220.0 / gc.radiusearthkm + 1.0)) R(isimp) = 1.0;
        double sfour = ss;
        double qzms24 = qzms2t;
        const double perige = (rp - 1.0) * gc.radiusearthkm;
        if (perige < 156.0) {
            sfour = perige - 78.0;
            if (perige < 98.0) sfour = 20.0;
            const double q = (120.0 - sfour) / gc.radiusearthkm;
            qzms24 = q * q * q * q;
            sfour = sfour / gc.radiusearthkm + 1.0;
        }
        const double pinvsq = 1.0 / posq;
        const double tsi = 1.0 / (ao - sfour);
        const double eta = ao * ecco * tsi;
        R(eta) = eta;
        const double etasq = eta * eta;
        const double eeta = ecco * eta;
        const double psisq = fabs(1.0 - etasq);
        const double coef = qzms24 * pow(tsi, 4.0);
        const double coef1 = coef / pow(psisq, 3.5);
        const double cc2 =
            coef1 * no *
            (ao * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
             0.375 * gc.j2 * tsi / psisq * R(con41) * (8.0 + 3.0 * etasq * (8.0 + etasq)));
        R(cc1) = bstar * cc2;
        double cc3 = 0.0;
        if (ecco > 1.0e-4) cc3 = -2.0 * coef * tsi * gc.j3oj2 * no * sinio / ecco;
        R(x1mth2) = 1.0 - cosio2;
        R(cc4) = 2.0 * no * coef1 * ao * omeosq *
                 (eta * (2.0 + 0.5 * etasq) + ecco * (0.5 + 2.0 * etasq) -
                  gc.j2 * tsi / (ao * psisq) *
                      (-3.0 * R(con41) * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                       0.75 * R(x1mth2) * (2.0 * etasq - eeta * (1.0 + etasq)) *
                           cos(2.0 * argpo)));
        R(cc5) = 2.0 * coef1 * ao * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
        const double cosio4 = cosio2 * cosio2;
        const double temp1 = 1.5 * gc.j2 * pinvsq * no;
        const double temp2 = 0.5 * temp1 * gc.j2 * pinvsq;
        const double temp3 = -0.46875 * gc.j4 * pinvsq * pinvsq * no;
        R(mdot) = no + 0.5 * temp1 * rteosq * R(con41) +
                  0.0625 * temp2 * rteosq * (13.0 - 78.0 * cosio2 + 137.0 * cosio4);
        R(argpdot) = -0.5 * temp1 * con42 +
                     0.0625 * temp2 * (7.0 - 114.0 * cosio2 + 395.0 * cosio4) +
                     temp3 * (3.0 - 36.0 * cosio2 + 49.0 * cosio4);
        const double xhdot1 = -temp1 * cosio;
        R(nodedot) = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * cosio2) +
                               2.0 * temp3 * (3.0 - 7.0 * cosio2)) * cosio;
        const double xpidot = R(argpdot) + R(nodedot);
        R(omgcof) = bstar * cc3 * cos(argpo);
        R(xmcof) = 0.0;
        if (ecco > 1.0e-4) R(xmcof) = -x2o3 * coef * bstar / eeta;
        R(nodecf) = 3.5 * omeosq * xhdot1 * R(cc1);
        R(t2cof) = 1.5 * R(cc1);
        if (fabs(cosio + 1.0) > 1.5e-12) {
            R(xlcof) = -0.25 * gc.j3oj2 * sinio * (3.0 + 5.0 * cosio) / (1.0 + cosio);
        } else {
            R(xlcof) = -0.25 * gc.j3oj2 * sinio * (3.0 + 5.0 * cosio) / 1.5e-12;
        }
        R(aycof) = -0.5 * gc.j3oj2 * sinio;
        const double delmotemp = 1.0 + eta * cos(mo);
        R(delmo) = delmotemp * delmotemp * delmotemp;
        R(sinmao) = sin(mo);
        R(x7thm1) = 7.0 * cosio2 - 1.0;

        if ((2 * kPi / no) >= 225.0) {
            R(method_d) = 1.0;
            R(isimp) = 1.0;
            deep_space_setup(rec, epoch, eccsq, xpidot);
        }

        if (R(isimp) != 1.0) {
            const double cc1 = R(cc1);
            const double cc1sq = cc1 * cc1;
            R(d2) = 4.0 * ao * tsi * cc1sq;
            const double temp = R(d2) * tsi * cc1 / 3.0;
            R(d3) = (17.0 * ao + sfour) * temp;
            R(d4) = 0.5 * temp * ao * tsi * (221.0 * ao + 31.0 * sfour) * cc1;
            R(t3cof) = R(d2) + 2.0 * cc1sq;
            R(t4cof) = 0.25 * (3.0 * R(d3) + cc1 * (12.0 * R(d2) + 10.0 * cc1sq));
            R(t5cof) = 0.2 * (3.0 * R(d4) + 12.0 * cc1 * R(d3) + 6.0 * R(d2) * R(d2) +
                              15.0 * cc1sq * (2.0 * R(d2) + cc1sq));
        }
    }
}

// Propagation-layout fields derived from a record (what kernel 2 reads per evaluation).
__device__ inline void derive_prop_fields(const SatRec &rec, double pf[PF_COUNT]) {
    pf[PF_JD0] = R(jdsatepoch);
    pf[PF_JDF0] = R(jdsatepochF);
    pf[PF_MO] = R(mo);
    pf[PF_MDOT] = R(mdot);
    pf[PF_ARGPO] = R(argpo);
    pf[PF_ARGPDOT] = R(argpdot);
    pf[PF_NODEO] = R(nodeo);
    pf[PF_NODEDOT] = R(nodedot);
    pf[PF_NODECF] = R(nodecf);
    pf[PF_CC1] = R(cc1);
    pf[PF_BC4] = R(bstar) * R(cc4);
    pf[PF_T2COF] = R(t2cof);
    pf[PF_NO] = R(no_unkozai);
    pf[PF_ECCO] = R(ecco);
    pf[PF_INCLO] = R(inclo);
    pf[PF_AOBASE] = pow(R(xke) / R(no_unkozai), 2.0 / 3.0);
    pf[PF_SINIO] = sin(R(inclo));
    pf[PF_COSIO] = cos(R(inclo));
    pf[PF_AYCOF] = R(aycof);
    pf[PF_XLCOF] = R(xlcof);
    pf[PF_CON41] = R(con41);
    pf[PF_X1MTH2] = R(x1mth2);
    pf[PF_X7THM1] = R(x7thm1);
    pf[PF_OMGCOF] = R(omgcof);
    pf[PF_ETA] = R(eta);
    pf[PF_XMCOF] = R(xmcof);
    pf[PF_DELMO] = R(delmo);
    pf[PF_D2] = R(d2);
    pf[PF_D3] = R(d3);
    pf[PF_D4] = R(d4);
    pf[PF_BC5] = R(bstar) * R(cc5);
    pf[PF_SINMAO] = R(sinmao);
    pf[PF_T3COF] = R(t3cof);
    pf[PF_T4COF] = R(t4cof);
    pf[PF_T5COF] = R(t5cof);
    pf[PF_DEDT] = R(dedt);
    pf[PF_DIDT] = R(didt);
    pf[PF_DMDT] = R(dmdt);
    pf[PF_DNODT] = R(dnodt);
    pf[PF_DOMDT] = R(domdt);
    pf[PF_GSTO] = R(gsto);
    pf[PF_E3] = R(e3);
    pf[PF_EE2] = R(ee2);
    pf[PF_SE2] = R(se2);
    pf[PF_SE3] = R(se3);
    pf[PF_SGH2] = R(sgh2);
    pf[PF_SGH3] = R(sgh3);
    pf[PF_SGH4] = R(sgh4);
    pf[PF_SH2] = R(sh2);
    pf[PF_SH3] = R(sh3);
    pf[PF_SI2] = R(si2);
    pf[PF_SI3] = R(si3);
    pf[PF_SL2] = R(sl2);
    pf[PF_SL3] = R(sl3);
    pf[PF_SL4] = R(sl4);
    pf[PF_XGH2] = R(xgh2);
    pf[PF_XGH3] = R(xgh3);
    pf[PF_XGH4] = R(xgh4);
    pf[PF_XH2] = R(xh2);
    pf[PF_XH3] = R(xh3);
    pf[PF_XI2] = R(xi2);
    pf[PF_XI3] = R(xi3);
    pf[PF_XL2] = R(xl2);
    pf[PF_XL3] = R(xl3);
    pf[PF_XL4] = R(xl4);
    pf[PF_ZMOL] = R(zmol);
    pf[PF_ZMOS] = R(zmos);
    pf[PF_XFACT] = R(xfact);
    pf[PF_XLAMO] = R(xlamo);
    pf[PF_DEL1] = R(del1);
    pf[PF_DEL2] = R(del2);
    pf[PF_DEL3] = R(del3);
    pf[PF_D2201] = R(d2201);
    pf[PF_D2211] = R(d2211);
    pf[PF_D3210] = R(d3210);
    pf[PF_D3222] = R(d3222);
    pf[PF_D4410] = R(d4410);
    pf[PF_D4422] = R(d4422);
    pf[PF_D5220] = R(d5220);
    pf[PF_D5232] = R(d5232);
    pf[PF_D5421] = R(d5421);
    pf[PF_D5433] = R(d5433);
}

__device__ inline int class_of(const SatRec &rec) {
    if (R(method_d) != 0.0) {
        const int irez = static_cast<int>(R(irez));
        return irez == 1 ? PB_CLS_DEEP_R1 : irez == 2 ? PB_CLS_DEEP_R2 : PB_CLS_DEEP_NR;
    }
    return (R(isimp) != 0.0) ? PB_CLS_NEAR_SIMP : PB_CLS_NEAR_FULL;
}

#undef R

}  // namespace pb

#endif
