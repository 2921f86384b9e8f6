/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the SGP4/SDP4 `Satellite::propagate` path.
 *
 * Plain-C restatement of the reference algorithm (gunvirranu/perturb wrapping Vallado's
 * SGP4 2020-07-13). Every function cites the reference file:line it follows. It keeps the
 * reference's floating-point operation ORDER expression by expression and is compiled
 * without FMA contraction, so on this toolchain it is bit-identical to the compiled
 * reference (oracle/_ref) -- tests/test_oracle.py pins that on SGP4-VER.TLE, the ISS
 * example, the published Vallado tcppver anchors and synthetic catalogues.
 *
 * Who may use it: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / reference arm.
 * The product (perturb_b200/, include/) never includes, links or calls anything here.
 */
#ifndef SGP4_ORACLE_H
#define SGP4_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_WGS72OLD = 0, ORC_WGS72 = 1, ORC_WGS84 = 2 };

/* Numeric content of one TLE in TLE units (deg, rev/day ...), i.e. what
 * perturb::TwoLineElement carries (include/perturb/tle.hpp:65-91). */
typedef struct orc_tle {
    double epoch_year;        /* two-digit year */
    double epoch_day_of_year; /* 1.0 .. 366.999 */
    double n_dot, n_ddot, b_star;
    double inclination, raan, eccentricity, arg_of_perigee, mean_anomaly, mean_motion;
} orc_tle;

/* The per-satellite record: the numeric subset of sgp4::elsetrec
 * (include/perturb/sgp4.hpp:89-134) that the propagate path reads or writes. */
typedef struct orc_sat {
    int error, isimp, irez, deep; /* deep: method == 'd' */
    int opsmode_a;                /* operationmode == 'a' */
    int grav;
    /* near earth */
    double aycof, con41, cc1, cc4, cc5, d2, d3, d4, delmo, eta, argpdot, omgcof, sinmao, t,
        t2cof, t3cof, t4cof, t5cof, x1mth2, x7thm1, mdot, nodedot, xlcof, xmcof, nodecf;
    /* deep space */
    double d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433, dedt, del1,
        del2, del3, didt, dmdt, dnodt, domdt, e3, ee2, peo, pgho, pho, pinco, plo, se2, se3,
        sgh2, sgh3, sgh4, sh2, sh3, si2, si3, sl2, sl3, sl4, gsto, xfact, xgh2, xgh3, xgh4,
        xh2, xh3, xi2, xi3, xl2, xl3, xl4, xlamo, zmol, zmos, atime, xli, xni;
    double a, altp, alta, jdsatepoch, jdsatepochF, nddot, ndot, bstar, inclo, nodeo, ecco,
        argpo, mo, no_kozai, no_unkozai;
    double am, em, im, Om, om, mm, nm;
    double tumin, mus, radiusearthkm, xke, j2, j3, j4, j3oj2;
    /* diagnostics of the last orc_sgp4 call (not part of elsetrec): Kepler iterations done and
     * the last Newton correction. kepler_iters == 10 with |kepler_last| >= 1e-12 means the
     * reference's loop (src/sgp4.cpp:1973-1983) ran out of iterations without converging: the
     * state it then returns is a function of rounding noise, not of the orbit. */
    double kepler_last;
    double cond; /* am / pl = 1 / (1 - el2): how much the short-period step amplifies rounding */
    int kepler_iters, pad_;
} orc_sat;

/* getgravconst, src/sgp4.cpp:2101-2157 */
void orc_gravity(int grav, orc_sat *s);
/* gstime_SGP4, src/sgp4.cpp:2506-2525 */
double orc_gstime(double jdut1);
/* days2mdhms_SGP4 / jday_SGP4, src/sgp4.cpp:3161-3192, 3100-3121 */
void orc_days2mdhms(int year, double days, int *mon, int *day, int *hr, int *minute, double *sec);
void orc_jday(int year, int mon, int day, int hr, int minute, double sec, double *jd, double *jdfrac);

/* sgp4init (+initl, dscom, dpper('y'), dsinit, trailing sgp4(0)), src/sgp4.cpp:1383-1680 */
void orc_sgp4init(orc_sat *s, int grav, char opsmode, double epoch, double bstar, double ndot,
                  double nddot, double ecco, double argpo, double inclo, double mo,
                  double no_kozai, double nodeo);
/* sgp4 (+dspace, dpper('n')), src/sgp4.cpp:1769-2065. Returns 1 when error == 0.
 * r, v are written only for codes 0 and 6, as in the reference. */
int orc_sgp4(orc_sat *s, double tsince, double r[3], double v[3]);

/* Satellite(const TwoLineElement&, GravModel), src/perturb.cpp:135-186 (opsmode 'i' there;
 * exposed so the verification catalogue can be run with 'a'). */
void orc_sat_from_tle(orc_sat *s, const orc_tle *tle, int grav, char opsmode);
/* The numeric part of twoline2rv's post-processing, src/sgp4.cpp:2336-2369, 2472-2474. */
void orc_sat_from_vallado_fields(orc_sat *s, const orc_tle *tle, int grav, char opsmode);

/* Satellite::propagate / propagate_from_epoch, src/perturb.cpp:228-242. Return Sgp4Error. */
int orc_propagate_mins(orc_sat *s, double mins, double r[3], double v[3]);
int orc_propagate_jd(orc_sat *s, double jd, double jd_frac, double r[3], double v[3]);

/* TLE text -> orc_tle. flavor 0: perturb's TwoLineElement::parse conventions
 * (src/tle.cpp:41-173); flavor 1: Vallado's twoline2rv conventions (src/sgp4.cpp:2231-2338).
 * Returns 0 on success, else a TLEParseError-like code (1 space, 2 format, 3 value,
 * 4 checksum) for flavor 0; flavor 1 never fails on 69-column input. */
int orc_parse_tle(const char *line1, const char *line2, int flavor, orc_tle *out);

/* n satellites x m times; sat-major outputs pos/vel [n][m][3], err [n][m]; NaN for codes
 * 1-4. use_jd: ta/tb are (jd, jd_frac) pairs, else ta are minutes from epoch. Each satellite
 * is walked through the grid in order on a private copy; satellites are split in contiguous
 * slices over nthreads. Returns wall seconds. */
double orc_propagate_grid(const orc_sat *sats, long n, const double *ta, const double *tb,
                          long m, int use_jd, int nthreads, double *pos, double *vel, int *err);
/* Same, and also reports per evaluation whether Kepler's equation converged (1) or the loop
 * ran out of iterations (0), and the conditioning figure `cond`; both [n][m], may be NULL. */
double orc_propagate_grid_diag(const orc_sat *sats, long n, const double *ta, const double *tb,
                               long m, int use_jd, int nthreads, double *pos, double *vel,
                               int *err, unsigned char *converged, double *cond);
/* Same, also returning what sgp4() leaves in the record per call: elems [n][m][7] = am, em,
 * im, Om, om, mm, nm; NaN where the call returned before storing them. May be NULL. */
double orc_propagate_grid_elems(const orc_sat *sats, long n, const double *ta, const double *tb,
                                long m, int use_jd, int nthreads, double *pos, double *vel,
                                int *err, double *elems);

/* ClassicalOrbitalElements(StateVector, GravModel) -- src/perturb.cpp:119-131, i.e.
 * rv2coe_SGP4 (src/sgp4.cpp:2863-3064) with newtonnu_SGP4 (:2749-2803) and angle_SGP4
 * (:2659-2681). out[11] = p, a, ecc, incl, omega, argp, nu, m, arglat, truelon, lonper. */
void orc_rv2coe(const double r[3], const double v[3], int grav, double out[11]);

int orc_sizeof_sat(void);

#ifdef __cplusplus
}
#endif
#endif
