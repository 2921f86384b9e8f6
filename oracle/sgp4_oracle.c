/* TEST INFRASTRUCTURE ONLY -- see sgp4_oracle.h. CPU oracle ("port") of the reference's
 * SGP4/SDP4 path, restated in plain C99. Parity pinned bit-exactly against the compiled
 * reference (oracle/_ref) by tests/test_oracle.py; never used by the product path.
 *
 * Floating-point contract: each expression keeps the reference's evaluation order, libm
 * calls are the same glibc functions with the same arguments, and the file is built with
 * -ffp-contract=off, so results are bit-identical to the reference built with
 * `-O3 -DNDEBUG` (no -march => no FMA).
 */
#define _POSIX_C_SOURCE 200809L
#include "sgp4_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PI 3.14159265358979323846 /* src/sgp4.cpp:69 */
static const double TWOPI = 2.0 * PI;
static const double X2O3 = 2.0 / 3.0;

/* lunar / solar constants shared by dscom, dsinit and dpper
 * (src/sgp4.cpp:259-262, 462-469, 764-765) */
static const double ZNS = 1.19459e-5, ZES = 0.01675, ZNL = 1.5835218e-4, ZEL = 0.05490;
static const double RPTIM = 4.37526908801129966e-3; /* src/sgp4.cpp:760, 1044 */

int orc_sizeof_sat(void) { return (int) sizeof(orc_sat); }

/* ---------------------------------------------------------------------------------------
 * getgravconst -- src/sgp4.cpp:2101-2157
 * ------------------------------------------------------------------------------------- */
void orc_gravity(int grav, orc_sat *s) {
    s->grav = grav;
    if (grav == ORC_WGS72OLD) {
        s->mus = 398600.79964;
        s->radiusearthkm = 6378.135;
        s->xke = 0.0743669161;
        s->tumin = 1.0 / s->xke;
        s->j2 = 0.001082616;
        s->j3 = -0.00000253881;
        s->j4 = -0.00000165597;
    } else if (grav == ORC_WGS84) {
        s->mus = 398600.5;
        s->radiusearthkm = 6378.137;
        s->xke = 60.0 / sqrt(s->radiusearthkm * s->radiusearthkm * s->radiusearthkm / s->mus);
        s->tumin = 1.0 / s->xke;
        s->j2 = 0.00108262998905;
        s->j3 = -0.00000253215306;
        s->j4 = -0.00000161098761;
    } else {
        s->mus = 398600.8;
        s->radiusearthkm = 6378.135;
        s->xke = 60.0 / sqrt(s->radiusearthkm * s->radiusearthkm * s->radiusearthkm / s->mus);
        s->tumin = 1.0 / s->xke;
        s->j2 = 0.001082616;
        s->j3 = -0.00000253881;
        s->j4 = -0.00000165597;
    }
    s->j3oj2 = s->j3 / s->j2;
}

/* gstime_SGP4 -- src/sgp4.cpp:2506-2525 */
double orc_gstime(double jdut1) {
    const double deg2rad = PI / 180.0;
    double tut1 = (jdut1 - 2451545.0) / 36525.0;
    double g = -6.2e-6 * tut1 * tut1 * tut1 + 0.093104 * tut1 * tut1 +
               (876600.0 * 3600 + 8640184.812866) * tut1 + 67310.54841;
    g = fmod(g * deg2rad / 240.0, TWOPI);
    if (g < 0.0) g += TWOPI;
    return g;
}

/* days2mdhms_SGP4 -- src/sgp4.cpp:3161-3192 */
void orc_days2mdhms(int year, double days, int *mon, int *day, int *hr, int *minute,
                    double *sec) {
    int mlen[13] = {0, 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int doy = (int) floor(days);
    if (year % 4 == 0) mlen[2] = 29;
    int m = 1, acc = 0;
    while (doy > acc + mlen[m] && m < 12) {
        acc += mlen[m];
        m++;
    }
    *mon = m;
    *day = doy - acc;
    double rem = (days - doy) * 24.0;
    *hr = (int) floor(rem);
    rem = (rem - *hr) * 60.0;
    *minute = (int) floor(rem);
    *sec = (rem - *minute) * 60.0;
}

/* jday_SGP4 -- src/sgp4.cpp:3100-3121 */
void orc_jday(int year, int mon, int day, int hr, int minute, double sec, double *jd,
              double *jdfrac) {
    *jd = 367.0 * year - floor((7 * (year + floor((mon + 9) / 12.0))) * 0.25) +
          floor(275 * mon / 9.0) + day + 1721013.5;
    *jdfrac = (sec + minute * 60.0 + hr * 3600.0) / 86400.0;
    if (fabs(*jdfrac) > 1.0) {
        double whole = floor(*jdfrac);
        *jd = *jd + whole;
        *jdfrac = *jdfrac - whole;
    }
}

/* ---------------------------------------------------------------------------------------
 * dpper -- src/sgp4.cpp:235-368. `apply`==0 is the init=='y' call (periodics are computed
 * at t=0 and nothing is applied, :296); `apply`==1 is the propagation call.
 * ------------------------------------------------------------------------------------- */
static void lunisolar_periodics(const orc_sat *s, double t, int apply, double *ep,
                                double *inclp, double *nodep, double *argpp, double *mp) {
    double zm, zf, sinzf, f2, f3;

    zm = s->zmos + ZNS * t;
    if (!apply) zm = s->zmos;
    zf = zm + 2.0 * ZES * sin(zm);
    sinzf = sin(zf);
    f2 = 0.5 * sinzf * sinzf - 0.25;
    f3 = -0.5 * sinzf * cos(zf);
    const double ses = s->se2 * f2 + s->se3 * f3;
    const double sis = s->si2 * f2 + s->si3 * f3;
    const double sls = s->sl2 * f2 + s->sl3 * f3 + s->sl4 * sinzf;
    const double sghs = s->sgh2 * f2 + s->sgh3 * f3 + s->sgh4 * sinzf;
    const double shs = s->sh2 * f2 + s->sh3 * f3;

    zm = s->zmol + ZNL * t;
    if (!apply) zm = s->zmol;
    zf = zm + 2.0 * ZEL * sin(zm);
    sinzf = sin(zf);
    f2 = 0.5 * sinzf * sinzf - 0.25;
    f3 = -0.5 * sinzf * cos(zf);
    const double sel = s->ee2 * f2 + s->e3 * f3;
    const double sil = s->xi2 * f2 + s->xi3 * f3;
    const double sll = s->xl2 * f2 + s->xl3 * f3 + s->xl4 * sinzf;
    const double sghl = s->xgh2 * f2 + s->xgh3 * f3 + s->xgh4 * sinzf;
    const double shll = s->xh2 * f2 + s->xh3 * f3;

    double pe = ses + sel;
    double pinc = sis + sil;
    double pl = sls + sll;
    double pgh = sghs + sghl;
    double ph = shs + shll;
    if (!apply) return;

    pe = pe - s->peo;
    pinc = pinc - s->pinco;
    pl = pl - s->plo;
    pgh = pgh - s->pgho;
    ph = ph - s->pho;
    *inclp = *inclp + pinc;
    *ep = *ep + pe;
    const double sinip = sin(*inclp);
    const double cosip = cos(*inclp);

    if (*inclp >= 0.2) { /* :317 direct application */
        ph = ph / sinip;
        pgh = pgh - cosip * ph;
        *argpp = *argpp + pgh;
        *nodep = *nodep + ph;
        *mp = *mp + pl;
    } else { /* :325-364 Lyddane modification */
        const double sinop = sin(*nodep);
        const double cosop = cos(*nodep);
        double alfdp = sinip * sinop;
        double betdp = sinip * cosop;
        const double dalf = ph * cosop + pinc * cosip * sinop;
        const double dbet = -ph * sinop + pinc * cosip * cosop;
        alfdp = alfdp + dalf;
        betdp = betdp + dbet;
        *nodep = fmod(*nodep, TWOPI);
        if (*nodep < 0.0 && s->opsmode_a) *nodep = *nodep + TWOPI;
        double xls = *mp + *argpp + cosip * *nodep;
        const double dls = pl + pgh - pinc * *nodep * sinip;
        xls = xls + dls;
        const double xnoh = *nodep;
        *nodep = atan2(alfdp, betdp);
        if (*nodep < 0.0 && s->opsmode_a) *nodep = *nodep + TWOPI;
        if (fabs(xnoh - *nodep) > PI) {
            if (*nodep < xnoh)
                *nodep = *nodep + TWOPI;
            else
                *nodep = *nodep - TWOPI;
        }
        *mp = *mp + pl;
        *argpp = xls - *mp - cosip * *nodep;
    }
}

/* ---------------------------------------------------------------------------------------
 * dscom -- src/sgp4.cpp:438-636. The reference runs one loop body twice (sun, then moon,
 * :525-602); here the body is a function returning the per-body terms.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    double s1, s2, s3, s4, s5, s6, s7;
    double z1, z2, z3, z11, z12, z13, z21, z22, z23, z31, z32, z33;
} body_terms;

typedef struct {
    double sinim, cosim, sinomm, cosomm, snodm, cnodm;
    double em, emsq, betasq, rtemsq, xnoi;
} orbit_geom;

static body_terms third_body(const orbit_geom *g, double zcosg, double zsing, double zcosi,
                             double zsini, double zcosh, double zsinh, double cc) {
    body_terms b;
    const double a1 = zcosg * zcosh + zsing * zcosi * zsinh;
    const double a3 = -zsing * zcosh + zcosg * zcosi * zsinh;
    const double a7 = -zcosg * zsinh + zsing * zcosi * zcosh;
    const double a8 = zsing * zsini;
    const double a9 = zsing * zsinh + zcosg * zcosi * zcosh;
    const double a10 = zcosg * zsini;
    const double a2 = g->cosim * a7 + g->sinim * a8;
    const double a4 = g->cosim * a9 + g->sinim * a10;
    const double a5 = -g->sinim * a7 + g->cosim * a8;
    const double a6 = -g->sinim * a9 + g->cosim * a10;

    const double x1 = a1 * g->cosomm + a2 * g->sinomm;
    const double x2 = a3 * g->cosomm + a4 * g->sinomm;
    const double x3 = -a1 * g->sinomm + a2 * g->cosomm;
    const double x4 = -a3 * g->sinomm + a4 * g->cosomm;
    const double x5 = a5 * g->sinomm;
    const double x6 = a6 * g->sinomm;
    const double x7 = a5 * g->cosomm;
    const double x8 = a6 * g->cosomm;
    const double emsq = g->emsq;

    b.z31 = 12.0 * x1 * x1 - 3.0 * x3 * x3;
    b.z32 = 24.0 * x1 * x2 - 6.0 * x3 * x4;
    b.z33 = 12.0 * x2 * x2 - 3.0 * x4 * x4;
    b.z1 = 3.0 * (a1 * a1 + a2 * a2) + b.z31 * emsq;
    b.z2 = 6.0 * (a1 * a3 + a2 * a4) + b.z32 * emsq;
    b.z3 = 3.0 * (a3 * a3 + a4 * a4) + b.z33 * emsq;
    b.z11 = -6.0 * a1 * a5 + emsq * (-24.0 * x1 * x7 - 6.0 * x3 * x5);
    b.z12 = -6.0 * (a1 * a6 + a3 * a5) +
            emsq * (-24.0 * (x2 * x7 + x1 * x8) - 6.0 * (x3 * x6 + x4 * x5));
    b.z13 = -6.0 * a3 * a6 + emsq * (-24.0 * x2 * x8 - 6.0 * x4 * x6);
    b.z21 = 6.0 * a2 * a5 + emsq * (24.0 * x1 * x5 - 6.0 * x3 * x7);
    b.z22 = 6.0 * (a4 * a5 + a2 * a6) +
            emsq * (24.0 * (x2 * x5 + x1 * x6) - 6.0 * (x4 * x7 + x3 * x8));
    b.z23 = 6.0 * a4 * a6 + emsq * (24.0 * x2 * x6 - 6.0 * x4 * x8);
    b.z1 = b.z1 + b.z1 + g->betasq * b.z31;
    b.z2 = b.z2 + b.z2 + g->betasq * b.z32;
    b.z3 = b.z3 + b.z3 + g->betasq * b.z33;
    b.s3 = cc * g->xnoi;
    b.s2 = -0.5 * b.s3 / g->rtemsq;
    b.s4 = b.s3 * g->rtemsq;
    b.s1 = -15.0 * g->em * b.s4;
    b.s5 = x1 * x3 + x2 * x4;
    b.s6 = x2 * x3 + x1 * x4;
    b.s7 = x2 * x4 - x1 * x3;
    return b;
}

static void deep_common(orc_sat *s, double epoch, double tc, orbit_geom *g, body_terms *sun,
                        body_terms *moon) {
    const double c1ss = 2.9864797e-6, c1l = 4.7968065e-7;
    const double zsinis = 0.39785416, zcosis = 0.91744867;
    const double zcosgs = 0.1945905, zsings = -0.98088458;

    g->snodm = sin(s->nodeo);
    g->cnodm = cos(s->nodeo);
    g->sinomm = sin(s->argpo);
    g->cosomm = cos(s->argpo);
    g->sinim = sin(s->inclo);
    g->cosim = cos(s->inclo);
    g->em = s->ecco;
    g->emsq = g->em * g->em;
    g->betasq = 1.0 - g->emsq;
    g->rtemsq = sqrt(g->betasq);
    g->xnoi = 1.0 / s->no_unkozai;

    s->peo = s->pinco = s->plo = s->pgho = s->pho = 0.0;
    const double day = epoch + 18261.5 + tc / 1440.0;
    const double xnodce = fmod(4.5236020 - 9.2422029e-4 * day, TWOPI);
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

    *sun = third_body(g, zcosgs, zsings, zcosis, zsinis, g->cnodm, g->snodm, c1ss);
    *moon = third_body(g, zcosgl, zsingl, zcosil, zsinil, zcoshl * g->cnodm + zsinhl * g->snodm,
                       g->snodm * zcoshl - g->cnodm * zsinhl, c1l);

    s->zmol = fmod(4.7199672 + 0.22997150 * day - gam, TWOPI);
    s->zmos = fmod(6.2565837 + 0.017201977 * day, TWOPI);

    /* solar periodic coefficients, :608-619 */
    s->se2 = 2.0 * sun->s1 * sun->s6;
    s->se3 = 2.0 * sun->s1 * sun->s7;
    s->si2 = 2.0 * sun->s2 * sun->z12;
    s->si3 = 2.0 * sun->s2 * (sun->z13 - sun->z11);
    s->sl2 = -2.0 * sun->s3 * sun->z2;
    s->sl3 = -2.0 * sun->s3 * (sun->z3 - sun->z1);
    s->sl4 = -2.0 * sun->s3 * (-21.0 - 9.0 * g->emsq) * ZES;
    s->sgh2 = 2.0 * sun->s4 * sun->z32;
    s->sgh3 = 2.0 * sun->s4 * (sun->z33 - sun->z31);
    s->sgh4 = -18.0 * sun->s4 * ZES;
    s->sh2 = -2.0 * sun->s2 * sun->z22;
    s->sh3 = -2.0 * sun->s2 * (sun->z23 - sun->z21);
    /* lunar periodic coefficients, :622-633 */
    s->ee2 = 2.0 * moon->s1 * moon->s6;
    s->e3 = 2.0 * moon->s1 * moon->s7;
    s->xi2 = 2.0 * moon->s2 * moon->z12;
    s->xi3 = 2.0 * moon->s2 * (moon->z13 - moon->z11);
    s->xl2 = -2.0 * moon->s3 * moon->z2;
    s->xl3 = -2.0 * moon->s3 * (moon->z3 - moon->z1);
    s->xl4 = -2.0 * moon->s3 * (-21.0 - 9.0 * g->emsq) * ZEL;
    s->xgh2 = 2.0 * moon->s4 * moon->z32;
    s->xgh3 = 2.0 * moon->s4 * (moon->z33 - moon->z31);
    s->xgh4 = -18.0 * moon->s4 * ZEL;
    s->xh2 = -2.0 * moon->s2 * moon->z22;
    s->xh3 = -2.0 * moon->s2 * (moon->z23 - moon->z21);
}

/* ---------------------------------------------------------------------------------------
 * dsinit -- src/sgp4.cpp:719-943. Called from sgp4init with t = tc = 0 and
 * em = ecco, inclm = inclo, nm = no_unkozai, argpm = nodem = mm = 0, so the "+ rate * t"
 * updates at :811-815 do not influence anything stored; they are kept for fidelity.
 * ------------------------------------------------------------------------------------- */
static void deep_init(orc_sat *s, const orbit_geom *g, const body_terms *sun,
                      const body_terms *moon, double eccsq, double xpidot) {
    const double q22 = 1.7891679e-6, q31 = 2.1460748e-6, q33 = 2.2123015e-7;
    const double root22 = 1.7891679e-6, root44 = 7.3636953e-9, root54 = 2.1765803e-9;
    const double root32 = 3.7393792e-7, root52 = 1.1428639e-7;
    const double t = 0.0, tc = 0.0;
    const double cosim = g->cosim, sinim = g->sinim;
    double em = g->em, emsq = g->emsq;
    double inclm = s->inclo, nm = s->no_unkozai;

    s->irez = 0;
    if (nm < 0.0052359877 && nm > 0.0034906585) s->irez = 1;
    if (nm >= 8.26e-3 && nm <= 9.24e-3 && em >= 0.5) s->irez = 2;

    /* solar secular rates */
    const double ses = sun->s1 * ZNS * sun->s5;
    const double sis = sun->s2 * ZNS * (sun->z11 + sun->z13);
    const double sls = -ZNS * sun->s3 * (sun->z1 + sun->z3 - 14.0 - 6.0 * emsq);
    const double sghs = sun->s4 * ZNS * (sun->z31 + sun->z33 - 6.0);
    double shs = -ZNS * sun->s2 * (sun->z21 + sun->z23);
    if (inclm < 5.2359877e-2 || inclm > PI - 5.2359877e-2) shs = 0.0;
    if (sinim != 0.0) shs = shs / sinim;
    const double sgs = sghs - cosim * shs;

    /* lunar secular rates */
    s->dedt = ses + moon->s1 * ZNL * moon->s5;
    s->didt = sis + moon->s2 * ZNL * (moon->z11 + moon->z13);
    s->dmdt = sls - ZNL * moon->s3 * (moon->z1 + moon->z3 - 14.0 - 6.0 * emsq);
    const double sghl = moon->s4 * ZNL * (moon->z31 + moon->z33 - 6.0);
    double shll = -ZNL * moon->s2 * (moon->z21 + moon->z23);
    if (inclm < 5.2359877e-2 || inclm > PI - 5.2359877e-2) shll = 0.0;
    s->domdt = sgs + sghl;
    s->dnodt = shs;
    if (sinim != 0.0) {
        s->domdt = s->domdt - cosim / sinim * shll;
        s->dnodt = s->dnodt + shll / sinim;
    }

    const double dndt = 0.0;
    const double theta = fmod(s->gsto + tc * RPTIM, TWOPI);
    em = em + s->dedt * t;
    inclm = inclm + s->didt * t;
    (void) inclm;

    if (s->irez == 0) return;
    const double aonv = pow(nm / s->xke, X2O3);

    if (s->irez == 2) { /* 12 h resonance, :831-915 */
        const double cosisq = cosim * cosim;
        const double emo = em;
        em = s->ecco;
        const double emsqo = emsq;
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
            if (em > 0.715)
                g520 = -5149.66 + 29936.92 * em - 54087.36 * emsq + 31324.56 * eoc;
            else
                g520 = 1464.74 - 4664.75 * em + 3763.64 * emsq;
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
        double temp = temp1 * root22;
        s->d2201 = temp * f220 * g201;
        s->d2211 = temp * f221 * g211;
        temp1 = temp1 * aonv;
        temp = temp1 * root32;
        s->d3210 = temp * f321 * g310;
        s->d3222 = temp * f322 * g322;
        temp1 = temp1 * aonv;
        temp = 2.0 * temp1 * root44;
        s->d4410 = temp * f441 * g410;
        s->d4422 = temp * f442 * g422;
        temp1 = temp1 * aonv;
        temp = temp1 * root52;
        s->d5220 = temp * f522 * g520;
        s->d5232 = temp * f523 * g532;
        temp = 2.0 * temp1 * root54;
        s->d5421 = temp * f542 * g521;
        s->d5433 = temp * f543 * g533;
        s->xlamo = fmod(s->mo + s->nodeo + s->nodeo - theta - theta, TWOPI);
        s->xfact = s->mdot + s->dmdt + 2.0 * (s->nodedot + s->dnodt - RPTIM) - s->no_unkozai;
        em = emo;
        emsq = emsqo;
    }

    if (s->irez == 1) { /* synchronous resonance, :918-933 */
        const double g200 = 1.0 + emsq * (-2.5 + 0.8125 * emsq);
        const double g310 = 1.0 + 2.0 * emsq;
        const double g300 = 1.0 + emsq * (-6.0 + 6.60937 * emsq);
        const double f220 = 0.75 * (1.0 + cosim) * (1.0 + cosim);
        const double f311 = 0.9375 * sinim * sinim * (1.0 + 3.0 * cosim) - 0.75 * (1.0 + cosim);
        double f330 = 1.0 + cosim;
        f330 = 1.875 * f330 * f330 * f330;
        s->del1 = 3.0 * nm * nm * aonv * aonv;
        s->del2 = 2.0 * s->del1 * f220 * g200 * q22;
        s->del3 = 3.0 * s->del1 * f330 * g300 * q33 * aonv;
        s->del1 = s->del1 * f311 * g310 * q31 * aonv;
        s->xlamo = fmod(s->mo + s->nodeo + s->argpo - theta, TWOPI);
        s->xfact = s->mdot + xpidot - RPTIM + s->dmdt + s->domdt + s->dnodt - s->no_unkozai;
    }

    /* integrator reset, :936-939 */
    s->xli = s->xlamo;
    s->xni = s->no_unkozai;
    s->atime = 0.0;
    (void) dndt;
}

/* ---------------------------------------------------------------------------------------
 * dspace -- src/sgp4.cpp:1018-1166
 * ------------------------------------------------------------------------------------- */
static void deep_secular(orc_sat *s, double t, double *em, double *argpm, double *inclm,
                         double *mm, double *nodem, double *nm) {
    const double fasx2 = 0.13130908, fasx4 = 2.8843198, fasx6 = 0.37448087;
    const double g22 = 5.7686396, g32 = 0.95240898, g44 = 1.8014998, g52 = 1.0508330,
                 g54 = 4.4108898;
    const double stepp = 720.0, stepn = -720.0, step2 = 259200.0;
    const double tc = t;

    double dndt = 0.0;
    const double theta = fmod(s->gsto + tc * RPTIM, TWOPI);
    *em = *em + s->dedt * t;
    *inclm = *inclm + s->didt * t;
    *argpm = *argpm + s->domdt * t;
    *nodem = *nodem + s->dnodt * t;
    *mm = *mm + s->dmdt * t;

    if (s->irez == 0) return;

    if (s->atime == 0.0 || t * s->atime <= 0.0 || fabs(t) < fabs(s->atime)) {
        s->atime = 0.0;
        s->xni = s->no_unkozai;
        s->xli = s->xlamo;
    }
    const double delt = (t > 0.0) ? stepp : stepn;
    double ft = 0.0, xndt = 0.0, xldot = 0.0, xnddt = 0.0;
    for (;;) {
        if (s->irez != 2) {
            xndt = s->del1 * sin(s->xli - fasx2) + s->del2 * sin(2.0 * (s->xli - fasx4)) +
                   s->del3 * sin(3.0 * (s->xli - fasx6));
            xldot = s->xni + s->xfact;
            xnddt = s->del1 * cos(s->xli - fasx2) + 2.0 * s->del2 * cos(2.0 * (s->xli - fasx4)) +
                    3.0 * s->del3 * cos(3.0 * (s->xli - fasx6));
            xnddt = xnddt * xldot;
        } else {
            const double xomi = s->argpo + s->argpdot * s->atime;
            const double x2omi = xomi + xomi;
            const double x2li = s->xli + s->xli;
            xndt = s->d2201 * sin(x2omi + s->xli - g22) + s->d2211 * sin(s->xli - g22) +
                   s->d3210 * sin(xomi + s->xli - g32) + s->d3222 * sin(-xomi + s->xli - g32) +
                   s->d4410 * sin(x2omi + x2li - g44) + s->d4422 * sin(x2li - g44) +
                   s->d5220 * sin(xomi + s->xli - g52) + s->d5232 * sin(-xomi + s->xli - g52) +
                   s->d5421 * sin(xomi + x2li - g54) + s->d5433 * sin(-xomi + x2li - g54);
            xldot = s->xni + s->xfact;
            xnddt = s->d2201 * cos(x2omi + s->xli - g22) + s->d2211 * cos(s->xli - g22) +
                    s->d3210 * cos(xomi + s->xli - g32) + s->d3222 * cos(-xomi + s->xli - g32) +
                    s->d5220 * cos(xomi + s->xli - g52) + s->d5232 * cos(-xomi + s->xli - g52) +
                    2.0 * (s->d4410 * cos(x2omi + x2li - g44) + s->d4422 * cos(x2li - g44) +
                           s->d5421 * cos(xomi + x2li - g54) + s->d5433 * cos(-xomi + x2li - g54));
            xnddt = xnddt * xldot;
        }
        if (fabs(t - s->atime) >= stepp) {
            s->xli = s->xli + xldot * delt + xndt * step2;
            s->xni = s->xni + xndt * delt + xnddt * step2;
            s->atime = s->atime + delt;
        } else {
            ft = t - s->atime;
            break;
        }
    }
    *nm = s->xni + xndt * ft + xnddt * ft * ft * 0.5;
    const double xl = s->xli + xldot * ft + xndt * ft * ft * 0.5;
    if (s->irez != 1) {
        *mm = xl - 2.0 * *nodem + 2.0 * theta;
        dndt = *nm - s->no_unkozai;
    } else {
        *mm = xl - *nodem - *argpm + theta;
        dndt = *nm - s->no_unkozai;
    }
    *nm = s->no_unkozai + dndt;
}

/* ---------------------------------------------------------------------------------------
 * sgp4 -- src/sgp4.cpp:1769-2065
 * ------------------------------------------------------------------------------------- */
int orc_sgp4(orc_sat *s, double tsince, double r[3], double v[3]) {
    const double temp4 = 1.5e-12;
    const double vkmpersec = s->radiusearthkm * s->xke / 60.0;
    double temp;

    s->t = tsince;
    s->error = 0;

    /* secular gravity and atmospheric drag, :1805-1835 */
    const double xmdf = s->mo + s->mdot * s->t;
    const double argpdf = s->argpo + s->argpdot * s->t;
    const double nodedf = s->nodeo + s->nodedot * s->t;
    double argpm = argpdf;
    double mm = xmdf;
    const double t2 = s->t * s->t;
    double nodem = nodedf + s->nodecf * t2;
    double tempa = 1.0 - s->cc1 * s->t;
    double tempe = s->bstar * s->cc4 * s->t;
    double templ = s->t2cof * t2;

    if (s->isimp != 1) {
        const double delomg = s->omgcof * s->t;
        const double delmtemp = 1.0 + s->eta * cos(xmdf);
        const double delm = s->xmcof * (delmtemp * delmtemp * delmtemp - s->delmo);
        temp = delomg + delm;
        mm = xmdf + temp;
        argpm = argpdf - temp;
        const double t3 = t2 * s->t;
        const double t4 = t3 * s->t;
        tempa = tempa - s->d2 * t2 - s->d3 * t3 - s->d4 * t4;
        tempe = tempe + s->bstar * s->cc5 * (sin(mm) - s->sinmao);
        templ = templ + s->t3cof * t3 + t4 * (s->t4cof + s->t * s->t5cof);
    }

    double nm = s->no_unkozai;
    double em = s->ecco;
    double inclm = s->inclo;
    if (s->deep) deep_secular(s, s->t, &em, &argpm, &inclm, &mm, &nodem, &nm);

    if (nm <= 0.0) {
        s->error = 2;
        return 0;
    }
    const double am = pow(s->xke / nm, X2O3) * tempa * tempa;
    nm = s->xke / pow(am, 1.5);
    em = em - tempe;
    if (em >= 1.0 || em < -0.001) {
        s->error = 1;
        return 0;
    }
    if (em < 1.0e-6) em = 1.0e-6;
    mm = mm + s->no_unkozai * templ;
    double xlm = mm + argpm + nodem;

    nodem = fmod(nodem, TWOPI);
    argpm = fmod(argpm, TWOPI);
    xlm = fmod(xlm, TWOPI);
    mm = fmod(xlm - argpm - nodem, TWOPI);

    s->am = am;
    s->em = em;
    s->im = inclm;
    s->Om = nodem;
    s->om = argpm;
    s->mm = mm;
    s->nm = nm;

    const double sinim = sin(inclm);
    const double cosim = cos(inclm);

    /* lunar-solar periodics, :1908-1958 */
    double ep = em, xincp = inclm, argpp = argpm, nodep = nodem, mp = mm;
    double sinip = sinim, cosip = cosim;
    if (s->deep) {
        lunisolar_periodics(s, s->t, 1, &ep, &xincp, &nodep, &argpp, &mp);
        if (xincp < 0.0) {
            xincp = -xincp;
            nodep = nodep + PI;
            argpp = argpp - PI;
        }
        if (ep < 0.0 || ep > 1.0) {
            s->error = 3;
            return 0;
        }
        sinip = sin(xincp);
        cosip = cos(xincp);
        s->aycof = -0.5 * s->j3oj2 * sinip;
        if (fabs(cosip + 1.0) > 1.5e-12)
            s->xlcof = -0.25 * s->j3oj2 * sinip * (3.0 + 5.0 * cosip) / (1.0 + cosip);
        else
            s->xlcof = -0.25 * s->j3oj2 * sinip * (3.0 + 5.0 * cosip) / temp4;
    }
    const double axnl = ep * cos(argpp);
    temp = 1.0 / (am * (1.0 - ep * ep));
    const double aynl = ep * sin(argpp) + temp * s->aycof;
    const double xl = mp + argpp + nodep + temp * s->xlcof * axnl;

    /* Kepler's equation, :1965-1983 */
    const double u = fmod(xl - nodep, TWOPI);
    double eo1 = u, tem5 = 9999.9, sineo1 = 1, coseo1 = 1;
    int ktr = 1;
    while (fabs(tem5) >= 1.0e-12 && ktr <= 10) {
        sineo1 = sin(eo1);
        coseo1 = cos(eo1);
        tem5 = 1.0 - coseo1 * axnl - sineo1 * aynl;
        tem5 = (u - aynl * coseo1 + axnl * sineo1 - eo1) / tem5;
        if (fabs(tem5) >= 0.95) tem5 = tem5 > 0.0 ? 0.95 : -0.95;
        eo1 = eo1 + tem5;
        ktr = ktr + 1;
    }
    s->kepler_iters = ktr - 1;
    s->kepler_last = tem5;

    /* short-period preliminaries, :1986-2011 */
    const double ecose = axnl * coseo1 + aynl * sineo1;
    const double esine = axnl * sineo1 - aynl * coseo1;
    const double el2 = axnl * axnl + aynl * aynl;
    const double pl = am * (1.0 - el2);
    s->cond = am / pl;
    if (pl < 0.0) {
        s->error = 4;
        return 0;
    }
    const double rl = am * (1.0 - ecose);
    const double rdotl = sqrt(am) * esine / rl;
    const double rvdotl = sqrt(pl) / rl;
    const double betal = sqrt(1.0 - el2);
    temp = esine / (1.0 + betal);
    const double sinu = am / rl * (sineo1 - aynl - axnl * temp);
    const double cosu = am / rl * (coseo1 - axnl + aynl * temp);
    double su = atan2(sinu, cosu);
    const double sin2u = (cosu + cosu) * sinu;
    const double cos2u = 1.0 - 2.0 * sinu * sinu;
    temp = 1.0 / pl;
    const double temp1 = 0.5 * s->j2 * temp;
    const double temp2 = temp1 * temp;

    if (s->deep) { /* :2014-2020 */
        const double cosisq = cosip * cosip;
        s->con41 = 3.0 * cosisq - 1.0;
        s->x1mth2 = 1.0 - cosisq;
        s->x7thm1 = 7.0 * cosisq - 1.0;
    }
    const double mrt =
        rl * (1.0 - 1.5 * temp2 * betal * s->con41) + 0.5 * temp1 * s->x1mth2 * cos2u;
    su = su - 0.25 * temp2 * s->x7thm1 * sin2u;
    const double xnode = nodep + 1.5 * temp2 * cosip * sin2u;
    const double xinc = xincp + 1.5 * temp2 * cosip * sinip * cos2u;
    const double mvt = rdotl - nm * temp1 * s->x1mth2 * sin2u / s->xke;
    const double rvdot = rvdotl + nm * temp1 * (s->x1mth2 * cos2u + 1.5 * s->con41) / s->xke;

    /* orientation vectors and state, :2031-2052 */
    const double sinsu = sin(su), cossu = cos(su);
    const double snod = sin(xnode), cnod = cos(xnode);
    const double sini = sin(xinc), cosi = cos(xinc);
    const double xmx = -snod * cosi;
    const double xmy = cnod * cosi;
    const double ux = xmx * sinsu + cnod * cossu;
    const double uy = xmy * sinsu + snod * cossu;
    const double uz = sini * sinsu;
    const double vx = xmx * cossu - cnod * sinsu;
    const double vy = xmy * cossu - snod * sinsu;
    const double vz = sini * cossu;

    r[0] = (mrt * ux) * s->radiusearthkm;
    r[1] = (mrt * uy) * s->radiusearthkm;
    r[2] = (mrt * uz) * s->radiusearthkm;
    v[0] = (mvt * ux + rvdot * vx) * vkmpersec;
    v[1] = (mvt * uy + rvdot * vy) * vkmpersec;
    v[2] = (mvt * uz + rvdot * vz) * vkmpersec;

    if (mrt < 1.0) { /* :2056-2061 */
        s->error = 6;
        return 0;
    }
    return 1;
}

/* ---------------------------------------------------------------------------------------
 * sgp4init -- src/sgp4.cpp:1383-1680 (initl inlined from :1221-1299)
 * ------------------------------------------------------------------------------------- */
void orc_sgp4init(orc_sat *s, int grav, char opsmode, double epoch, double bstar, double ndot,
                  double nddot, double ecco, double argpo, double inclo, double mo,
                  double no_kozai, double nodeo) {
    const double jd0 = s->jdsatepoch, jdf0 = s->jdsatepochF;
    memset(s, 0, sizeof *s); /* :1415-1444 zeroing */
    s->jdsatepoch = jd0;
    s->jdsatepochF = jdf0;
    orc_gravity(grav, s);
    s->opsmode_a = (opsmode == 'a');
    s->bstar = bstar;
    s->ndot = ndot;
    s->nddot = nddot;
    s->ecco = ecco;
    s->argpo = argpo;
    s->inclo = inclo;
    s->mo = mo;
    s->no_kozai = no_kozai;
    s->nodeo = nodeo;

    const double ss = 78.0 / s->radiusearthkm + 1.0;
    const double qzms2ttemp = (120.0 - 78.0) / s->radiusearthkm;
    const double qzms2t = qzms2ttemp * qzms2ttemp * qzms2ttemp * qzms2ttemp;

    /* initl */
    const double eccsq = s->ecco * s->ecco;
    const double omeosq = 1.0 - eccsq;
    const double rteosq = sqrt(omeosq);
    const double cosio = cos(s->inclo);
    const double cosio2 = cosio * cosio;
    const double ak = pow(s->xke / s->no_kozai, X2O3);
    const double d1 = 0.75 * s->j2 * (3.0 * cosio2 - 1.0) / (rteosq * omeosq);
    double del = d1 / (ak * ak);
    const double adel = ak * (1.0 - del * del - del * (1.0 / 3.0 + 134.0 * del * del / 81.0));
    del = d1 / (adel * adel);
    s->no_unkozai = s->no_kozai / (1.0 + del);
    const double ao = pow(s->xke / (s->no_unkozai), X2O3);
    const double sinio = sin(s->inclo);
    const double po = ao * omeosq;
    const double con42 = 1.0 - 5.0 * cosio2;
    s->con41 = -con42 - cosio2 - cosio2;
    const double posq = po * po;
    const double rp = ao * (1.0 - s->ecco);
    s->deep = 0;
    s->gsto = orc_gstime(epoch + 2433281.5);

    s->a = pow(s->no_unkozai * s->tumin, (-2.0 / 3.0));
    s->alta = s->a * (1.0 + s->ecco) - 1.0;
    s->altp = s->a * (1.0 - s->ecco) - 1.0;
    s->error = 0;

    if (omeosq >= 0.0 || s->no_unkozai >= 0.0) {
        s->isimp = 0;
        if (rp < (220.0 / s->radiusearthkm + 1.0)) s->isimp = 1;
        double sfour = ss;
        double qzms24 = qzms2t;
        const double perige = (rp - 1.0) * s->radiusearthkm;
        if (perige < 156.0) {
            sfour = perige - 78.0;
            if (perige < 98.0) sfour = 20.0;
            const double q = (120.0 - sfour) / s->radiusearthkm;
            qzms24 = q * q * q * q;
            sfour = sfour / s->radiusearthkm + 1.0;
        }
        const double pinvsq = 1.0 / posq;
        const double tsi = 1.0 / (ao - sfour);
        s->eta = ao * s->ecco * tsi;
        const double etasq = s->eta * s->eta;
        const double eeta = s->ecco * s->eta;
        const double psisq = fabs(1.0 - etasq);
        const double coef = qzms24 * pow(tsi, 4.0);
        const double coef1 = coef / pow(psisq, 3.5);
        const double cc2 =
            coef1 * s->no_unkozai *
            (ao * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
             0.375 * s->j2 * tsi / psisq * s->con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)));
        s->cc1 = s->bstar * cc2;
        double cc3 = 0.0;
        if (s->ecco > 1.0e-4)
            cc3 = -2.0 * coef * tsi * s->j3oj2 * s->no_unkozai * sinio / s->ecco;
        s->x1mth2 = 1.0 - cosio2;
        s->cc4 = 2.0 * s->no_unkozai * coef1 * ao * omeosq *
                 (s->eta * (2.0 + 0.5 * etasq) + s->ecco * (0.5 + 2.0 * etasq) -
                  s->j2 * tsi / (ao * psisq) *
                      (-3.0 * s->con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                       0.75 * s->x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) *
                           cos(2.0 * s->argpo)));
        s->cc5 = 2.0 * coef1 * ao * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
        const double cosio4 = cosio2 * cosio2;
        const double temp1 = 1.5 * s->j2 * pinvsq * s->no_unkozai;
        const double temp2 = 0.5 * temp1 * s->j2 * pinvsq;
        const double temp3 = -0.46875 * s->j4 * pinvsq * pinvsq * s->no_unkozai;
        s->mdot = s->no_unkozai + 0.5 * temp1 * rteosq * s->con41 +
                  0.0625 * temp2 * rteosq * (13.0 - 78.0 * cosio2 + 137.0 * cosio4);
        s->argpdot = -0.5 * temp1 * con42 +
                     0.0625 * temp2 * (7.0 - 114.0 * cosio2 + 395.0 * cosio4) +
                     temp3 * (3.0 - 36.0 * cosio2 + 49.0 * cosio4);
        const double xhdot1 = -temp1 * cosio;
        s->nodedot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * cosio2) +
                               2.0 * temp3 * (3.0 - 7.0 * cosio2)) * cosio;
        const double xpidot = s->argpdot + s->nodedot;
        s->omgcof = s->bstar * cc3 * cos(s->argpo);
        s->xmcof = 0.0;
        if (s->ecco > 1.0e-4) s->xmcof = -X2O3 * coef * s->bstar / eeta;
        s->nodecf = 3.5 * omeosq * xhdot1 * s->cc1;
        s->t2cof = 1.5 * s->cc1;
        if (fabs(cosio + 1.0) > 1.5e-12)
            s->xlcof = -0.25 * s->j3oj2 * sinio * (3.0 + 5.0 * cosio) / (1.0 + cosio);
        else
            s->xlcof = -0.25 * s->j3oj2 * sinio * (3.0 + 5.0 * cosio) / 1.5e-12;
        s->aycof = -0.5 * s->j3oj2 * sinio;
        const double delmotemp = 1.0 + s->eta * cos(s->mo);
        s->delmo = delmotemp * delmotemp * delmotemp;
        s->sinmao = sin(s->mo);
        s->x7thm1 = 7.0 * cosio2 - 1.0;

        if ((2 * PI / s->no_unkozai) >= 225.0) { /* deep space, :1590-1649 */
            s->deep = 1;
            s->isimp = 1;
            orbit_geom g;
            body_terms sun, moon;
            deep_common(s, epoch, 0.0, &g, &sun, &moon);
            double e_ = s->ecco, i_ = s->inclo, n_ = s->nodeo, a_ = s->argpo, m_ = s->mo;
            lunisolar_periodics(s, s->t, 0, &e_, &i_, &n_, &a_, &m_);
            deep_init(s, &g, &sun, &moon, eccsq, xpidot);
        }

        if (s->isimp != 1) { /* :1652-1667 */
            const double cc1sq = s->cc1 * s->cc1;
            s->d2 = 4.0 * ao * tsi * cc1sq;
            const double temp = s->d2 * tsi * s->cc1 / 3.0;
            s->d3 = (17.0 * ao + sfour) * temp;
            s->d4 = 0.5 * temp * ao * tsi * (221.0 * ao + 31.0 * sfour) * s->cc1;
            s->t3cof = s->d2 + 2.0 * cc1sq;
            s->t4cof = 0.25 * (3.0 * s->d3 + s->cc1 * (12.0 * s->d2 + 10.0 * cc1sq));
            s->t5cof = 0.2 * (3.0 * s->d4 + 12.0 * s->cc1 * s->d3 + 6.0 * s->d2 * s->d2 +
                              15.0 * cc1sq * (2.0 * s->d2 + cc1sq));
        }
    }
    double r[3], v[3];
    orc_sgp4(s, 0.0, r, v); /* :1673 */
}

/* ---------------------------------------------------------------------------------------
 * Satellite(const TwoLineElement&, GravModel) -- src/perturb.cpp:135-186
 * ------------------------------------------------------------------------------------- */
static void finish_from_fields(orc_sat *s, const orc_tle *tle, int grav, char opsmode,
                               double no_kozai, double ndot, double nddot) {
    const double deg2rad = PI / 180.0;
    const int yy = (int) tle->epoch_year;
    const int year = yy + (yy < 57 ? 2000 : 1900);
    int mon, day, hr, minute;
    double sec, jd, jdf;
    orc_days2mdhms(year, tle->epoch_day_of_year, &mon, &day, &hr, &minute, &sec);
    orc_jday(year, mon, day, hr, minute, sec, &jd, &jdf);
    memset(s, 0, sizeof *s);
    s->jdsatepoch = jd;
    s->jdsatepochF = jdf;
    const double epoch = (jd + jdf) - 2433281.5;
    orc_sgp4init(s, grav, opsmode, epoch, tle->b_star, ndot, nddot, tle->eccentricity,
                 tle->arg_of_perigee * deg2rad, tle->inclination * deg2rad,
                 tle->mean_anomaly * deg2rad, no_kozai, tle->raan * deg2rad);
}

void orc_sat_from_tle(orc_sat *s, const orc_tle *tle, int grav, char opsmode) {
    const double xp_dot_p = 1440.0 / (2 * PI); /* perturb.cpp:137 */
    finish_from_fields(s, tle, grav, opsmode, tle->mean_motion / xp_dot_p,
                       tle->n_dot / (xp_dot_p * 1440.0), tle->n_ddot / (xp_dot_p * 1440.0 * 1440));
}

/* twoline2rv numeric post-processing -- src/sgp4.cpp:2336-2349 (same arithmetic as above;
 * kept separate because the two parsers deliver n_ddot / b_star differently scaled). */
void orc_sat_from_vallado_fields(orc_sat *s, const orc_tle *tle, int grav, char opsmode) {
    const double xpdotp = 1440.0 / (2.0 * PI);
    finish_from_fields(s, tle, grav, opsmode, tle->mean_motion / xpdotp,
                       tle->n_dot / (xpdotp * 1440.0), tle->n_ddot / (xpdotp * 1440.0 * 1440));
}

/* Sgp4Error mapping -- src/perturb.cpp:24-29 */
static int map_error(int code) { return (code < 0 || code >= 8) ? 8 : code; }

/* Satellite::propagate_from_epoch -- src/perturb.cpp:228-234 */
int orc_propagate_mins(orc_sat *s, double mins, double r[3], double v[3]) {
    orc_sgp4(s, mins, r, v);
    return map_error(s->error);
}

/* Satellite::propagate -- src/perturb.cpp:236-242, with JulianDate::operator- (:80-83) */
int orc_propagate_jd(orc_sat *s, double jd, double jd_frac, double r[3], double v[3]) {
    const double delta_jd = (jd - s->jdsatepoch) + (jd_frac - s->jdsatepochF);
    const double mins = delta_jd * (24 * 60);
    return orc_propagate_mins(s, mins, r, v);
}

/* ---------------------------------------------------------------------------------------
 * TLE text -> numbers (fixed columns, strtod = what sscanf("%lf") does).
 * flavor 0: TwoLineElement::parse (src/tle.cpp:41-173): mantissa read as an integer-valued
 *           double, exponent reduced by 5 (:123-128), value = mant * pow(10, exp) (:136-137),
 *           eccentricity = int / 1e7 (:138).
 * flavor 1: twoline2rv (src/sgp4.cpp:2231-2338): implied decimal point inserted, value =
 *           0.mant * pow(10, exp), eccentricity read as ".ddddddd".
 * ------------------------------------------------------------------------------------- */
static double field_num(const char *line, int from, int len) {
    char buf[40];
    memcpy(buf, line + from, (size_t) len);
    buf[len] = '\0';
    return strtod(buf, NULL);
}

static int field_int(const char *line, int from, int len) {
    char buf[40];
    memcpy(buf, line + from, (size_t) len);
    buf[len] = '\0';
    return (int) strtol(buf, NULL, 10);
}

static double implied_point(const char *line, int sign_col, int len) {
    /* sign at sign_col, `len` digits after it, decimal point implied before the digits */
    char buf[40];
    int k = 0;
    if (line[sign_col] == '-') buf[k++] = '-';
    buf[k++] = '.';
    for (int i = 1; i <= len; ++i) {
        const char c = line[sign_col + i];
        buf[k++] = (c == ' ') ? '0' : c;
    }
    buf[k] = '\0';
    return strtod(buf, NULL);
}

static unsigned checksum69(const char *line) { /* src/tle.cpp:25-36 */
    unsigned sum = 0;
    for (int i = 0; i < 68; ++i) {
        if (line[i] >= '0' && line[i] <= '9') sum += (unsigned) (line[i] - '0');
        if (line[i] == '-') sum += 1u;
    }
    return sum % 10u;
}

int orc_parse_tle(const char *l1, const char *l2, int flavor, orc_tle *out) {
    memset(out, 0, sizeof *out);
    if (strlen(l1) < 69 || strlen(l2) < 69) return 2;
    if (flavor == 0) {
        static const int sp1[] = {2, 9, 18, 33, 44, 53, 62, 64};
        static const int sp2[] = {2, 8, 17, 26, 34, 43, 52};
        for (unsigned i = 0; i < sizeof sp1 / sizeof *sp1; ++i)
            if (l1[sp1[i] - 1] != ' ') return 1;
        for (unsigned i = 0; i < sizeof sp2 / sizeof *sp2; ++i)
            if (l2[sp2[i] - 1] != ' ') return 1;
    }
    out->epoch_year = field_int(l1, 18, 2);
    out->epoch_day_of_year = field_num(l1, 20, 12);
    out->n_dot = field_num(l1, 33, 10);
    const int nexp = field_int(l1, 50, 2);
    const int bexp = field_int(l1, 59, 2);
    if (flavor == 0) {
        out->n_ddot = field_num(l1, 44, 6) * pow(10.0, nexp - 5);
        out->b_star = field_num(l1, 53, 6) * pow(10.0, bexp - 5);
        out->eccentricity = (double) (unsigned long) field_int(l2, 26, 7) / 1.0e7;
    } else {
        out->n_ddot = implied_point(l1, 44, 5) * pow(10.0, nexp);
        out->b_star = implied_point(l1, 53, 5) * pow(10.0, bexp);
        out->eccentricity = implied_point(l2, 25, 7);
    }
    out->inclination = field_num(l2, 8, 8);
    out->raan = field_num(l2, 17, 8);
    out->arg_of_perigee = field_num(l2, 34, 8);
    out->mean_anomaly = field_num(l2, 43, 8);
    out->mean_motion = field_num(l2, 52, 11);
    if (flavor == 0) {
        int ok = (l1[0] == '1') && (l2[0] == '2');
        ok &= (l1[7] == 'U' || l1[7] == 'C' || l1[7] == 'S');
        ok &= (out->epoch_day_of_year >= 1.0 && out->epoch_day_of_year <= 366.0);
        ok &= (nexp - 5 > -15 && nexp - 5 < 10) && (bexp - 5 > -15 && bexp - 5 < 10);
        ok &= (l1[62] == '0');
        ok &= (memcmp(l1 + 2, l2 + 2, 5) == 0);
        ok &= (out->inclination >= 0.0 && out->inclination <= 180.0);
        ok &= (out->raan >= 0.0 && out->raan <= 360.0);
        ok &= (out->arg_of_perigee >= 0.0 && out->arg_of_perigee <= 360.0);
        ok &= (out->mean_anomaly >= 0.0 && out->mean_anomaly <= 360.0);
        if (!ok) return 3;
        if (checksum69(l1) != (unsigned) (l1[68] - '0')) return 4;
        if (checksum69(l2) != (unsigned) (l2[68] - '0')) return 4;
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------
 * Grid driver (the scalar loop of BASELINE.md section 3), pthreads over satellite slices.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    const orc_sat *sats;
    long lo, hi, m;
    const double *ta, *tb;
    int use_jd;
    double *pos, *vel;
    int *err;
    unsigned char *converged;
    double *cond;
    double *elems;
} grid_job;

static void *grid_worker(void *arg) {
    grid_job *j = (grid_job *) arg;
    for (long i = j->lo; i < j->hi; ++i) {
        orc_sat s = j->sats[i];
        for (long k = 0; k < j->m; ++k) {
            s.kepler_iters = 0;
            s.kepler_last = 0.0;
            s.cond = 0.0;
            if (j->elems) s.am = s.em = s.im = s.Om = s.om = s.mm = s.nm = NAN;
            double r[3] = {NAN, NAN, NAN}, v[3] = {NAN, NAN, NAN};
            int e;
            if (j->use_jd)
                e = orc_propagate_jd(&s, j->ta[k], j->tb[k], r, v);
            else
                e = orc_propagate_mins(&s, j->ta[k], r, v);
            const long o = i * j->m + k;
            if (j->pos) {
                for (int c = 0; c < 3; ++c) {
                    j->pos[o * 3 + c] = r[c];
                    j->vel[o * 3 + c] = v[c];
                }
            }
            if (j->err) j->err[o] = e;
            if (j->converged)
                j->converged[o] = !(s.kepler_iters >= 10 && fabs(s.kepler_last) >= 1.0e-12);
            if (j->cond) j->cond[o] = s.cond;
            if (j->elems) {
                double *q = j->elems + o * 7;
                q[0] = s.am; q[1] = s.em; q[2] = s.im; q[3] = s.Om; q[4] = s.om; q[5] = s.mm;
                q[6] = s.nm;
            }
        }
    }
    return NULL;
}

double orc_propagate_grid(const orc_sat *sats, long n, const double *ta, const double *tb,
                          long m, int use_jd, int nthreads, double *pos, double *vel, int *err) {
    return orc_propagate_grid_diag(sats, n, ta, tb, m, use_jd, nthreads, pos, vel, err, NULL,
                                   NULL);
}

static double grid_run(const orc_sat *sats, long n, const double *ta, const double *tb, long m,
                       int use_jd, int nthreads, double *pos, double *vel, int *err,
                       unsigned char *converged, double *cond, double *elems);

double orc_propagate_grid_diag(const orc_sat *sats, long n, const double *ta, const double *tb,
                               long m, int use_jd, int nthreads, double *pos, double *vel,
                               int *err, unsigned char *converged, double *cond) {
    return grid_run(sats, n, ta, tb, m, use_jd, nthreads, pos, vel, err, converged, cond, NULL);
}

double orc_propagate_grid_elems(const orc_sat *sats, long n, const double *ta, const double *tb,
                                long m, int use_jd, int nthreads, double *pos, double *vel,
                                int *err, double *elems) {
    return grid_run(sats, n, ta, tb, m, use_jd, nthreads, pos, vel, err, NULL, NULL, elems);
}

static double grid_run(const orc_sat *sats, long n, const double *ta, const double *tb, long m,
                       int use_jd, int nthreads, double *pos, double *vel, int *err,
                       unsigned char *converged, double *cond, double *elems) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[1024];
    grid_job jobs[1024];
    for (int t = 0; t < nthreads; ++t) {
        grid_job j = {sats, n * t / nthreads, n * (t + 1) / nthreads, m, ta, tb, use_jd, pos, vel, err,
                      converged, cond, elems};
        jobs[t] = j;
        if (nthreads == 1)
            grid_worker(&jobs[t]);
        else
            pthread_create(&th[t], NULL, grid_worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/* ---------------------------------------------------------------------------------------
 * rv2coe_SGP4 -- src/sgp4.cpp:2863-3064 (post-processing of a propagated state; "next" row
 * N4 of SURVEY.md section 8f). Helpers: mag/dot/cross (:2564-2633), angle_SGP4 (:2659-2681),
 * newtonnu_SGP4 (:2749-2803), asinh_SGP4 (:2705-2711).
 * ------------------------------------------------------------------------------------- */
static double vmag(const double x[3]) { return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }
static double vdot(const double x[3], const double y[3]) {
    return (x[0] * y[0] + x[1] * y[1] + x[2] * y[2]);
}
static double vsgn(double x) { return x < 0.0 ? -1.0 : 1.0; }

static double vangle(const double a[3], const double b[3]) {
    const double small = 0.00000001, undefined = 999999.1;
    const double ma = vmag(a), mb = vmag(b);
    if (ma * mb > small * small) {
        double temp = vdot(a, b) / (ma * mb);
        if (fabs(temp) > 1.0) temp = vsgn(temp) * 1.0;
        return acos(temp);
    }
    return undefined;
}

static void kepler_from_true_anomaly(double ecc, double nu, double *e0, double *m) {
    const double small = 0.00000001;
    *e0 = 999999.9;
    *m = 999999.9;
    if (fabs(ecc) < small) {
        *m = nu;
        *e0 = nu;
    } else if (ecc < 1.0 - small) {
        const double sine = (sqrt(1.0 - ecc * ecc) * sin(nu)) / (1.0 + ecc * cos(nu));
        const double cose = (ecc + cos(nu)) / (1.0 + ecc * cos(nu));
        *e0 = atan2(sine, cose);
        *m = *e0 - ecc * sin(*e0);
    } else if (ecc > 1.0 + small) {
        if ((ecc > 1.0) && (fabs(nu) + 0.00001 < PI - acos(1.0 / ecc))) {
            const double sine = (sqrt(ecc * ecc - 1.0) * sin(nu)) / (1.0 + ecc * cos(nu));
            *e0 = log(sine + sqrt(sine * sine + 1.0));
            *m = ecc * sinh(*e0) - *e0;
        }
    } else if (fabs(nu) < 168.0 * PI / 180.0) {
        *e0 = tan(nu * 0.5);
        *m = *e0 + (*e0 * *e0 * *e0) / 3.0;
    }
    if (ecc < 1.0) {
        *m = fmod(*m, 2.0 * PI);
        if (*m < 0.0) *m = *m + 2.0 * PI;
        *e0 = fmod(*e0, 2.0 * PI);
    }
}

void orc_rv2coe(const double r[3], const double v[3], int grav, double out[11]) {
    const double small = 0.00000001, undefined = 999999.1, infinite = 999999.9;
    const double halfpi = 0.5 * PI;
    orc_sat g;
    orc_gravity(grav, &g);
    const double mus = g.mus;
    double p, a, ecc, incl, omega, argp, nu, m = 0.0, arglat, truelon, lonper;

    const double magr = vmag(r), magv = vmag(v);
    const double hbar[3] = {r[1] * v[2] - r[2] * v[1], r[2] * v[0] - r[0] * v[2],
                            r[0] * v[1] - r[1] * v[0]};
    const double magh = vmag(hbar);
    if (magh > small) {
        const double nbar[3] = {-hbar[1], hbar[0], 0.0};
        const double magn = vmag(nbar);
        const double c1 = magv * magv - mus / magr;
        const double rdotv = vdot(r, v);
        double ebar[3];
        for (int i = 0; i <= 2; i++) ebar[i] = (c1 * r[i] - rdotv * v[i]) / mus;
        ecc = vmag(ebar);
        const double sme = (magv * magv * 0.5) - (mus / magr);
        a = (fabs(sme) > small) ? -mus / (2.0 * sme) : infinite;
        p = magh * magh / mus;
        const double hk = hbar[2] / magh;
        incl = acos(hk);

        int orbit = 1; /* 1 'ei', 2 'ce', 3 'ci', 4 'ee' */
        const int equatorial = (incl < small) | (fabs(incl - PI) < small);
        if (ecc < small)
            orbit = equatorial ? 2 : 3;
        else if (equatorial)
            orbit = 4;

        if (magn > small) {
            double temp = nbar[0] / magn;
            if (fabs(temp) > 1.0) temp = vsgn(temp);
            omega = acos(temp);
            if (nbar[1] < 0.0) omega = TWOPI - omega;
        } else
            omega = undefined;

        if (orbit == 1) {
            argp = vangle(nbar, ebar);
            if (ebar[2] < 0.0) argp = TWOPI - argp;
        } else
            argp = undefined;

        if (orbit == 1 || orbit == 4) {
            nu = vangle(ebar, r);
            if (rdotv < 0.0) nu = TWOPI - nu;
        } else
            nu = undefined;

        if (orbit == 3) {
            arglat = vangle(nbar, r);
            if (r[2] < 0.0) arglat = TWOPI - arglat;
            m = arglat;
        } else
            arglat = undefined;

        if (ecc > small && orbit == 4) {
            double temp = ebar[0] / ecc;
            if (fabs(temp) > 1.0) temp = vsgn(temp);
            lonper = acos(temp);
            if (ebar[1] < 0.0) lonper = TWOPI - lonper;
            if (incl > halfpi) lonper = TWOPI - lonper;
        } else
            lonper = undefined;

        if (magr > small && orbit == 2) {
            double temp = r[0] / magr;
            if (fabs(temp) > 1.0) temp = vsgn(temp);
            truelon = acos(temp);
            if (r[1] < 0.0) truelon = TWOPI - truelon;
            if (incl > halfpi) truelon = TWOPI - truelon;
            m = truelon;
        } else
            truelon = undefined;

        if (orbit == 1 || orbit == 4) {
            double e;
            kepler_from_true_anomaly(ecc, nu, &e, &m);
        }
    } else {
        p = a = ecc = incl = omega = argp = nu = m = arglat = truelon = lonper = undefined;
    }
    out[0] = p; out[1] = a; out[2] = ecc; out[3] = incl; out[4] = omega; out[5] = argp;
    out[6] = nu; out[7] = m; out[8] = arglat; out[9] = truelon; out[10] = lonper;
}
