// Deterministic synthetic TLE catalogues (include/perturb_synth.h).
// std::mt19937_64, u = (next() >> 11) * 2^-53; values rounded to TLE field precision by the
// text formatting itself, so every consumer parses identical numbers.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/perturb_synth.h"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kMu = 398600.8;      // km^3/s^2 (WGS72)
constexpr double kRe = 6378.135;      // km

struct Rng {
    std::mt19937_64 gen;
    explicit Rng(uint64_t seed) : gen(seed) {}
    double u() { return static_cast<double>(gen() >> 11) * (1.0 / 9007199254740992.0); }
    double uniform(double lo, double hi) { return lo + (hi - lo) * u(); }
    double log_uniform(double lo, double hi) {
        return std::exp(uniform(std::log(lo), std::log(hi)));
    }
    double normal() {  // Box-Muller (std::normal_distribution is implementation-defined)
        const double u1 = 1.0 - u();
        const double u2 = u();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * kPi * u2);
    }
};

struct Elements {
    double n;     // rev/day
    double e;
    double i;     // deg
    double raan, argp, ma;  // deg
    double bstar;
    double epoch_day;  // day of year 2026
};

double mean_motion_for_sma(double a_km) {
    return 86400.0 / (2.0 * kPi) * std::sqrt(kMu / (a_km * a_km * a_km));
}

void common_angles(Rng &r, Elements &el) {
    el.raan = r.uniform(0.0, 360.0);
    el.argp = r.uniform(0.0, 360.0);
    el.ma = r.uniform(0.0, 360.0);
    el.epoch_day = 274.0 + r.uniform(0.0, 3.0);
}

Elements make_leo(Rng &r) {
    Elements el;
    common_angles(r, el);
    const double pick = r.u();
    if (pick < 0.005) {  // perigee < 156 km: s4/qoms24 adjustment (src/sgp4.cpp:1523)
        el.e = r.uniform(0.002, 0.02);
        el.n = mean_motion_for_sma((kRe + r.uniform(100.0, 150.0)) / (1.0 - el.e));
    } else if (pick < 0.02) {  // perigee < 220 km: simplified drag (src/sgp4.cpp:1516)
        el.e = r.uniform(0.002, 0.02);
        el.n = mean_motion_for_sma((kRe + r.uniform(160.0, 212.0)) / (1.0 - el.e));
    } else {
        el.n = (r.u() < 0.02) ? r.uniform(11.3, 12.5) : r.uniform(12.5, 15.9);
        el.e = std::min(0.05, std::fabs(0.003 * r.normal()));
        if (r.u() < 0.05) el.e = r.uniform(0.0, 1.0e-4);  // drops cc3/xmcof (sgp4.cpp:1547)
        // keep the perigee of ordinary members above the simplified-drag limit
        const double a = std::cbrt(kMu / std::pow(el.n * 2.0 * kPi / 86400.0, 2.0));
        if (a * (1.0 - el.e) - kRe < 240.0) el.e = std::max(0.0, 1.0 - (kRe + 240.0) / a);
    }
    const double ip = r.u();
    if (ip < 0.40) {
        el.i = 53.0 + 0.2 * r.normal();
    } else if (ip < 0.65) {
        el.i = 97.6 + 0.5 * r.normal();
    } else {
        el.i = r.uniform(0.0, 180.0);
    }
    el.bstar = r.log_uniform(1.0e-6, 1.0e-3);
    if (r.u() < 0.02) el.bstar = -el.bstar;
    return el;
}

Elements make_meo(Rng &r) {  // deep space, non-resonant
    Elements el;
    common_angles(r, el);
    if (r.u() < 0.5) {
        el.n = 2.0057 + 0.0005 * r.normal();  // GPS-like; e < 0.5 keeps it out of irez 2
    } else {
        el.n = r.uniform(2.2, 6.3);
    }
    el.e = r.uniform(0.0, 0.02);
    el.i = r.uniform(50.0, 65.0);
    el.bstar = r.log_uniform(1.0e-6, 1.0e-4);
    return el;
}

Elements make_geo(Rng &r) {  // 24 h resonance, irez 1 (src/sgp4.cpp:773)
    Elements el;
    common_angles(r, el);
    el.n = r.uniform(0.99, 1.01);
    el.e = r.uniform(0.0, 0.01);
    el.i = r.uniform(0.0, 15.0);  // about 3/4 below the 0.2 rad Lyddane switch
    el.bstar = 0.0;
    return el;
}

Elements make_molniya(Rng &r) {  // 12 h resonance, irez 2 (src/sgp4.cpp:775)
    Elements el;
    common_angles(r, el);
    el.n = r.uniform(1.95, 2.06);
    el.e = r.uniform(0.60, 0.74);
    el.i = 63.4 + 1.5 * r.normal();
    el.argp = 270.0 + 10.0 * r.normal();
    el.bstar = r.log_uniform(1.0e-5, 1.0e-3);
    return el;
}

Elements make_gto(Rng &r) {
    Elements el;
    common_angles(r, el);
    el.n = r.uniform(2.2, 2.4);
    el.e = 0.73 + 0.005 * r.normal();
    el.i = r.uniform(5.0, 28.0);
    el.bstar = r.log_uniform(1.0e-5, 1.0e-3);
    return el;
}

Elements make_mega(Rng &r) {
    static const double shells[5][3] = {  // inclination, mean motion, cumulative share
        {53.0, 15.06, 0.40}, {53.2, 15.10, 0.60}, {70.0, 14.98, 0.75},
        {97.6, 15.03, 0.90}, {43.0, 15.14, 1.00}};
    Elements el;
    common_angles(r, el);
    const double pick = r.u();
    int s = 0;
    while (s < 4 && pick >= shells[s][2]) ++s;
    el.i = shells[s][0] + 0.01 * r.normal();
    el.n = shells[s][1] + 0.002 * r.normal();
    el.e = r.log_uniform(5.0e-5, 2.0e-3);
    el.bstar = r.log_uniform(1.0e-5, 5.0e-4);
    return el;
}

double wrap360(double x) {
    x = std::fmod(x, 360.0);
    if (x < 0.0) x += 360.0;
    if (x >= 359.99995) x = 0.0;  // would print as 360.0000
    return x;
}

// "[sign]ddddd[sign]d": 0.ddddd x 10^exp
void exp_field(double v, char out[32]) {
    char sign = ' ';
    if (v < 0.0) {
        sign = '-';
        v = -v;
    }
    int ex = 0;
    long mant = 0;
    if (v > 0.0) {
        ex = static_cast<int>(std::floor(std::log10(v))) + 1;
        mant = std::lround(v / std::pow(10.0, ex) * 1.0e5);
        if (mant >= 100000) {
            mant = 10000;
            ex += 1;
        }
        if (ex > 9 || ex < -9) {
            mant = 0;
            ex = 0;
        }
    }
    std::snprintf(out, 32, "%c%05ld%c%1d", sign, mant % 100000, ex < 0 ? '-' : '+', std::abs(ex) % 10);
    if (mant == 0) std::memcpy(out, " 00000-0", 9);
}

char checksum(const char *line) {
    unsigned sum = 0;
    for (int i = 0; i < 68; ++i) {
        if (line[i] >= '0' && line[i] <= '9') sum += static_cast<unsigned>(line[i] - '0');
        if (line[i] == '-') sum += 1;
    }
    return static_cast<char>('0' + sum % 10);
}

void format_tle(const Elements &el, size_t index, char *l1, char *l2) {
    const unsigned satnum = static_cast<unsigned>(index % 100000);
    char bstar[32];
    exp_field(el.bstar, bstar);
    double incl = std::min(std::max(el.i, 0.0), 180.0);
    double ecc = std::min(std::max(el.e, 0.0), 0.9999999);
    long ecc7 = std::lround(ecc * 1.0e7);
    if (ecc7 > 9999999) ecc7 = 9999999;
    char buf[96];
    std::snprintf(buf, sizeof buf, "1 %05uU 26001A   26%012.8f  .00000000  00000-0 %s 0  9990",
                  satnum, el.epoch_day, bstar);
    std::memcpy(l1, buf, 69);
    l1[68] = checksum(l1);
    l1[69] = '\0';
    std::snprintf(buf, sizeof buf, "2 %05u %8.4f %8.4f %07ld %8.4f %8.4f %11.8f    10", satnum,
                  incl, wrap360(el.raan), ecc7, wrap360(el.argp), wrap360(el.ma), el.n);
    std::memcpy(l2, buf, 69);
    l2[68] = checksum(l2);
    l2[69] = '\0';
}

}  // namespace

extern "C" int perturb_synth_catalog(int config, size_t n, uint64_t seed, char *lines1,
                                     char *lines2) {
    if (lines1 == nullptr || lines2 == nullptr) return -1;
    if (config < PERTURB_SYNTH_LEO || config > PERTURB_SYNTH_MEGA) return -1;
    if (seed == 0) seed = 0xB2000000ULL + static_cast<uint64_t>(config);
    Rng rng(seed);
    std::vector<Elements> els(n);
    for (size_t i = 0; i < n; ++i) {
        switch (config) {
            case PERTURB_SYNTH_LEO: els[i] = make_leo(rng); break;
            case PERTURB_SYNTH_MIXED: {
                const double p = static_cast<double>(i) / static_cast<double>(n ? n : 1);
                els[i] = p < 0.70   ? make_leo(rng)
                         : p < 0.80 ? make_meo(rng)
                         : p < 0.90 ? make_geo(rng)
                         : p < 0.95 ? make_molniya(rng)
                                    : make_gto(rng);
                break;
            }
            case PERTURB_SYNTH_DEEP:
                els[i] = (i < n / 2) ? make_geo(rng) : make_molniya(rng);
                break;
            default: els[i] = make_mega(rng); break;
        }
    }
    if (config == PERTURB_SYNTH_MIXED) {  // shuffled input order: exercises the sort path
        for (size_t i = n; i > 1; --i) {
            const size_t j = static_cast<size_t>(rng.u() * static_cast<double>(i));
            std::swap(els[i - 1], els[std::min(j, i - 1)]);
        }
    }
    for (size_t i = 0; i < n; ++i) {
        format_tle(els[i], i, lines1 + i * PERTURB_SYNTH_LINE_STRIDE,
                   lines2 + i * PERTURB_SYNTH_LINE_STRIDE);
    }
    return 0;
}

extern "C" int perturb_synth_time_grid(size_t m, double *jd, double *jd_frac) {
    if (jd == nullptr || jd_frac == nullptr) return -1;
    // 2026 day 274.0 is JD 2461314.5 (jday_SGP4 of 2026-10-01 00:00); base = that + 3 days.
    for (size_t k = 0; k < m; ++k) {
        jd[k] = 2461317.5;
        jd_frac[k] = static_cast<double>(k) / 1440.0;
    }
    return 0;
}
