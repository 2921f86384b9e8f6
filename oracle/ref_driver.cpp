// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or shipped with the product.
//
// Thin extern "C" driver around the UNMODIFIED reference library (gunvirranu/perturb).
// It is compiled against the sources where they lie under /root/reference (see
// oracle/Makefile, target `_ref`) into oracle/_ref/libperturb_ref.so. Nothing of the
// reference is copied into this repository: this file only *calls* the reference's own
// public API so tests and the bench's CPU arm can drive it through ctypes:
//
//   Satellite::from_tle(char*, char*, GravModel)        src/perturb.cpp:189-205
//   Satellite::Satellite(const TwoLineElement&, Grav)   src/perturb.cpp:135-186
//   TwoLineElement::parse(const char*, const char*)     src/tle.cpp:41-173
//   sgp4::twoline2rv(..., 'v', ...) verification mode   src/sgp4.cpp:2201-2475
//   Satellite::propagate(JulianDate, StateVector&)      src/perturb.cpp:236-242
//   Satellite::propagate_from_epoch(double, StateVec&)  src/perturb.cpp:228-234
//   Satellite::last_error()/epoch()                     src/perturb.cpp:220-226
//
// Who may use it: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
// `--impl reference` arm. See oracle/README.md.

#include <perturb/perturb.hpp>
#include <perturb/sgp4.hpp>
#include <perturb/tle.hpp>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

using perturb::GravModel;
using perturb::JulianDate;
using perturb::Satellite;
using perturb::StateVector;
using perturb::TwoLineElement;
namespace sgp4 = perturb::sgp4;

namespace {

GravModel to_grav(int g) {
    switch (g) {
        case 0: return GravModel::WGS72_OLD;
        case 2: return GravModel::WGS84;
        default: return GravModel::WGS72;
    }
}

sgp4::gravconsttype to_gravconst(int g) {
    switch (g) {
        case 0: return sgp4::wgs72old;
        case 2: return sgp4::wgs84;
        default: return sgp4::wgs72;
    }
}

// Plain mirror of TwoLineElement's numeric members (no padding surprises over ctypes).
struct RefTle {
    char catalog_number[8];
    double epoch_year;
    double epoch_day_of_year;
    double n_dot;
    double n_ddot;
    double b_star;
    double inclination;
    double raan;
    double eccentricity;
    double arg_of_perigee;
    double mean_anomaly;
    double mean_motion;
    double revolution_number;
    double element_set_number;
    double ephemeris_type;
    double classification;
    double line_1_checksum;
    double line_2_checksum;
};

void fill_nan(StateVector &sv) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int k = 0; k < 3; ++k) {
        sv.position[k] = nan;
        sv.velocity[k] = nan;
    }
}

}  // namespace

extern "C" {

int ref_sizeof_satellite(void) { return static_cast<int>(sizeof(Satellite)); }

void *ref_alloc(long n) {
    return new std::vector<Satellite>(
        static_cast<std::size_t>(n), Satellite(sgp4::elsetrec {})
    );
}

void ref_free(void *h) { delete static_cast<std::vector<Satellite> *>(h); }

static Satellite &sat_at(void *h, long i) {
    return (*static_cast<std::vector<Satellite> *>(h))[static_cast<std::size_t>(i)];
}

// `Satellite::from_tle` (Vallado's parser, opsmode 'i'). Lines are copied because the
// reference mutates them. Returns last_error() as int.
int ref_init_from_tle(void *h, long i, const char *l1, const char *l2, int grav) {
    char b1[140], b2[140];
    std::memset(b1, 0, sizeof b1);
    std::memset(b2, 0, sizeof b2);
    std::strncpy(b1, l1, 130);
    std::strncpy(b2, l2, 130);
    sat_at(h, i) = Satellite::from_tle(b1, b2, to_grav(grav));
    return static_cast<int>(sat_at(h, i).last_error());
}

// `TwoLineElement::parse`; returns the TLEParseError as int.
int ref_parse_tle(const char *l1, const char *l2, RefTle *out) {
    TwoLineElement tle {};
    const auto err = tle.parse(l1, l2);
    std::memset(out, 0, sizeof *out);
    std::memcpy(out->catalog_number, tle.catalog_number, 6);
    out->epoch_year = tle.epoch_year;
    out->epoch_day_of_year = tle.epoch_day_of_year;
    out->n_dot = tle.n_dot;
    out->n_ddot = tle.n_ddot;
    out->b_star = tle.b_star;
    out->inclination = tle.inclination;
    out->raan = tle.raan;
    out->eccentricity = tle.eccentricity;
    out->arg_of_perigee = tle.arg_of_perigee;
    out->mean_anomaly = tle.mean_anomaly;
    out->mean_motion = tle.mean_motion;
    out->revolution_number = static_cast<double>(tle.revolution_number);
    out->element_set_number = tle.element_set_number;
    out->ephemeris_type = tle.ephemeris_type;
    out->classification = tle.classification;
    out->line_1_checksum = tle.line_1_checksum;
    out->line_2_checksum = tle.line_2_checksum;
    return static_cast<int>(err);
}

// `Satellite(const TwoLineElement&, GravModel)` from the numeric members only.
int ref_init_from_parsed(void *h, long i, const RefTle *in, int grav) {
    TwoLineElement tle {};
    std::memcpy(tle.catalog_number, in->catalog_number, 6);
    tle.classification = static_cast<char>(in->classification);
    tle.epoch_year = static_cast<unsigned int>(in->epoch_year);
    tle.epoch_day_of_year = in->epoch_day_of_year;
    tle.n_dot = in->n_dot;
    tle.n_ddot = in->n_ddot;
    tle.b_star = in->b_star;
    tle.ephemeris_type = static_cast<unsigned char>(in->ephemeris_type);
    tle.element_set_number = static_cast<unsigned int>(in->element_set_number);
    tle.inclination = in->inclination;
    tle.raan = in->raan;
    tle.eccentricity = in->eccentricity;
    tle.arg_of_perigee = in->arg_of_perigee;
    tle.mean_anomaly = in->mean_anomaly;
    tle.mean_motion = in->mean_motion;
    tle.revolution_number = static_cast<unsigned long>(in->revolution_number);
    sat_at(h, i) = Satellite(tle, to_grav(grav));
    return static_cast<int>(sat_at(h, i).last_error());
}

// Verification-mode construction exactly as tests/test_perturb.cpp:28-50 does it
// (`twoline2rv('v','e',opsmode,grav)`); also returns the grid in the line-2 tail.
int ref_init_verif(
    void *h, long i, const char *l1, const char *l2, char opsmode, int grav,
    double *startmfe, double *stopmfe, double *deltamin
) {
    char b1[140], b2[140];
    std::memset(b1, 0, sizeof b1);
    std::memset(b2, 0, sizeof b2);
    std::strncpy(b1, l1, 130);
    std::strncpy(b2, l2, 130);
    sgp4::elsetrec rec {};
    *startmfe = *stopmfe = *deltamin = 0.0;
#ifdef PERTURB_SGP4_ENABLE_DEBUG
    sgp4::twoline2rv(
        b1, b2, 'v', 'e', opsmode, to_gravconst(grav), *startmfe, *stopmfe, *deltamin, rec
    );
#else
    (void) opsmode;
    (void) to_gravconst;
    rec.error = 7;
#endif
    sat_at(h, i) = Satellite(rec);
    return static_cast<int>(sat_at(h, i).last_error());
}

// Raw `sgp4init` on already-converted elements (isolates kernel 1 for parity).
int ref_init_raw(
    void *h, long i, int grav, char opsmode, double jdsatepoch, double jdsatepochF,
    double bstar, double ndot, double nddot, double ecco, double argpo, double inclo,
    double mo, double no_kozai, double nodeo
) {
    sgp4::elsetrec rec {};
    std::strcpy(rec.satnum, "0");
    rec.jdsatepoch = jdsatepoch;
    rec.jdsatepochF = jdsatepochF;
    sgp4::sgp4init(
        to_gravconst(grav), opsmode, rec.satnum, (jdsatepoch + jdsatepochF) - 2433281.5,
        bstar, ndot, nddot, ecco, argpo, inclo, mo, no_kozai, nodeo, rec
    );
    sat_at(h, i) = Satellite(rec);
    return static_cast<int>(sat_at(h, i).last_error());
}

int ref_last_error(void *h, long i) { return static_cast<int>(sat_at(h, i).last_error()); }

void ref_epoch(void *h, long i, double *jd, double *jd_frac) {
    const JulianDate e = sat_at(h, i).epoch();
    *jd = e.jd;
    *jd_frac = e.jd_frac;
}

// `propagate_from_epoch` on satellite i (mutates it, like the reference does).
int ref_propagate_mins(void *h, long i, double mins, double *r, double *v) {
    StateVector sv;
    fill_nan(sv);
    const int err = static_cast<int>(sat_at(h, i).propagate_from_epoch(mins, sv));
    for (int k = 0; k < 3; ++k) {
        r[k] = sv.position[k];
        v[k] = sv.velocity[k];
    }
    return err;
}

// `propagate(JulianDate)` on satellite i.
int ref_propagate_jd(void *h, long i, double jd, double jd_frac, double *r, double *v) {
    StateVector sv;
    fill_nan(sv);
    const int err = static_cast<int>(sat_at(h, i).propagate(JulianDate(jd, jd_frac), sv));
    for (int k = 0; k < 3; ++k) {
        r[k] = sv.position[k];
        v[k] = sv.velocity[k];
    }
    return err;
}

// Field dump for init parity. Order is fixed; see oracle/ref.py FIELD_NAMES.
int ref_dump_fields(void *h, long i, double *out) {
    const sgp4::elsetrec &s = sat_at(h, i).sat_rec;
    int k = 0;
#define F(x) out[k++] = static_cast<double>(s.x)
    F(error); F(isimp); F(irez);
    out[k++] = (s.method == 'd') ? 1.0 : 0.0;
    F(aycof); F(con41); F(cc1); F(cc4); F(cc5); F(d2); F(d3); F(d4); F(delmo); F(eta);
    F(argpdot); F(omgcof); F(sinmao); F(t2cof); F(t3cof); F(t4cof); F(t5cof); F(x1mth2);
    F(x7thm1); F(mdot); F(nodedot); F(xlcof); F(xmcof); F(nodecf);
    F(d2201); F(d2211); F(d3210); F(d3222); F(d4410); F(d4422); F(d5220); F(d5232);
    F(d5421); F(d5433); F(dedt); F(del1); F(del2); F(del3); F(didt); F(dmdt); F(dnodt);
    F(domdt); F(e3); F(ee2); F(peo); F(pgho); F(pho); F(pinco); F(plo); F(se2); F(se3);
    F(sgh2); F(sgh3); F(sgh4); F(sh2); F(sh3); F(si2); F(si3); F(sl2); F(sl3); F(sl4);
    F(gsto); F(xfact); F(xgh2); F(xgh3); F(xgh4); F(xh2); F(xh3); F(xi2); F(xi3); F(xl2);
    F(xl3); F(xl4); F(xlamo); F(zmol); F(zmos); F(atime); F(xli); F(xni);
    F(a); F(altp); F(alta); F(jdsatepoch); F(jdsatepochF); F(nddot); F(ndot); F(bstar);
    F(inclo); F(nodeo); F(ecco); F(argpo); F(mo); F(no_kozai); F(no_unkozai);
    F(am); F(em); F(im); F(Om); F(om); F(mm); F(nm);
    F(tumin); F(mus); F(radiusearthkm); F(xke); F(j2); F(j3); F(j4); F(j3oj2);
#undef F
    return k;
}

// Grid evaluation: for every satellite i in [0,n) and every time k in [0,m):
//   use_jd != 0: propagate(JulianDate(ta[k], tb[k]))   else: propagate_from_epoch(ta[k])
// Outputs sat-major: pos/vel [n][m][3], err [n][m]. States for codes 1-4 stay NaN (the
// reference leaves the caller's StateVector untouched, src/sgp4.cpp:1865,1878,1943,1995).
// Each satellite is propagated on a private copy walked through the grid in order, one
// contiguous satellite slice per thread.
// Returns wall seconds of the propagation loop.
double ref_propagate_grid(
    void *h, long n, const double *ta, const double *tb, long m, int use_jd, int nthreads,
    double *pos, double *vel, int *err
) {
    if (nthreads < 1) nthreads = 1;
    auto work = [&](long lo, long hi) {
        for (long i = lo; i < hi; ++i) {
            Satellite sat = sat_at(h, i);
            for (long k = 0; k < m; ++k) {
                StateVector sv;
                fill_nan(sv);
                perturb::Sgp4Error e;
                if (use_jd) {
                    e = sat.propagate(JulianDate(ta[k], tb[k]), sv);
                } else {
                    e = sat.propagate_from_epoch(ta[k], sv);
                }
                const long o = i * m + k;
                if (pos) {
                    for (int c = 0; c < 3; ++c) {
                        pos[o * 3 + c] = sv.position[c];
                        vel[o * 3 + c] = sv.velocity[c];
                    }
                }
                if (err) err[o] = static_cast<int>(e);
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        work(0, n);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) {
            const long lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            pool.emplace_back(work, lo, hi);
        }
        for (auto &th : pool) th.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// `ClassicalOrbitalElements(StateVector, GravModel)` (src/perturb.cpp:119-131) for `count`
// states; r, v [count][3]; out [count][11] in the struct's member order.
void ref_rv2coe(const double *r, const double *v, long count, int grav, double *out) {
    for (long e = 0; e < count; ++e) {
        StateVector sv;
        for (int c = 0; c < 3; ++c) {
            sv.position[c] = r[e * 3 + c];
            sv.velocity[c] = v[e * 3 + c];
        }
        const perturb::ClassicalOrbitalElements coe(sv, to_grav(grav));
        double *o = out + e * 11;
        o[0] = coe.semilatus_rectum;
        o[1] = coe.semimajor_axis;
        o[2] = coe.eccentricity;
        o[3] = coe.inclination;
        o[4] = coe.raan;
        o[5] = coe.arg_of_perigee;
        o[6] = coe.true_anomaly;
        o[7] = coe.mean_anomaly;
        o[8] = coe.arg_of_latitude;
        o[9] = coe.true_longitude;
        o[10] = coe.longitude_of_periapsis;
    }
}

int ref_hardware_threads(void) {
    const unsigned c = std::thread::hardware_concurrency();
    return c ? static_cast<int>(c) : 1;
}

}  // extern "C"
