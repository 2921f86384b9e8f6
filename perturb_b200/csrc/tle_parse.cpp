// Host-side fixed-column TLE reader feeding the batch packer (perturb_gpu_parse_tles).
//
// The reference has two readers: perturb's own TwoLineElement::parse (src/tle.cpp:41-173,
// sscanf based) and Vallado's twoline2rv (src/sgp4.cpp:2231-2338, in-place string fix-ups +
// sscanf). A TLE is a fixed-column card image, so this reader slices columns directly and
// converts each slice with strtod -- the same correctly-rounded conversion sscanf("%lf")
// performs -- and then applies the numeric conventions of whichever reference reader the
// caller wants to be bit-compatible with (`flavor`).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "../../include/perturb_gpu.h"

namespace {

constexpr int kLineLen = 69;  // perturb::TLE_LINE_LEN, include/perturb/tle.hpp:33

struct Slice {
    const char *p;
    int len;
};

double to_double(Slice s) {
    char buf[32];
    const int n = s.len < 31 ? s.len : 31;
    std::memcpy(buf, s.p, static_cast<std::size_t>(n));
    buf[n] = '\0';
    return std::strtod(buf, nullptr);
}

long to_long(Slice s) {
    char buf[32];
    const int n = s.len < 31 ? s.len : 31;
    std::memcpy(buf, s.p, static_cast<std::size_t>(n));
    buf[n] = '\0';
    return std::strtol(buf, nullptr, 10);
}

// "[sign]ddddd" with the decimal point implied in front of the digits: value 0.ddddd
double implied_decimal(const char *sign_col, int digits) {
    char buf[32];
    int k = 0;
    if (*sign_col == '-') buf[k++] = '-';
    buf[k++] = '.';
    for (int i = 1; i <= digits; ++i) buf[k++] = (sign_col[i] == ' ') ? '0' : sign_col[i];
    buf[k] = '\0';
    return std::strtod(buf, nullptr);
}

unsigned line_checksum(const char *line) {  // src/tle.cpp:25-36
    unsigned sum = 0U;
    for (int i = 0; i < kLineLen - 1; ++i) {
        const char c = line[i];
        if (c >= '0' && c <= '9') sum += static_cast<unsigned>(c - '0');
        if (c == '-') sum += 1U;
    }
    return sum % 10U;
}

bool has_nul(const char *line) { return std::memchr(line, '\0', kLineLen) != nullptr; }

// Returns a perturb::TLEParseError value (0 NONE, 1 SHOULD_BE_SPACE, 2 INVALID_FORMAT,
// 3 INVALID_VALUE, 4 CHECKSUM_MISMATCH), checked in that order like src/tle.cpp:39-44 says.
int parse_one(const char *l1, const char *l2, int flavor, perturb_gpu_tle &t) {
    if (has_nul(l1) || has_nul(l2)) return 2;  // shorter than TLE_LINE_LEN
    if (flavor == 0) {
        static const int spaces1[] = {2, 9, 18, 33, 44, 53, 62, 64};  // src/tle.cpp:43
        static const int spaces2[] = {2, 8, 17, 26, 34, 43, 52};      // src/tle.cpp:49
        for (int c : spaces1) {
            if (l1[c - 1] != ' ') return 1;
        }
        for (int c : spaces2) {
            if (l2[c - 1] != ' ') return 1;
        }
    }

    t.epoch_year = static_cast<double>(to_long({l1 + 18, 2}));
    t.epoch_day_of_year = to_double({l1 + 20, 12});
    t.n_dot = to_double({l1 + 33, 10});
    const long nddot_exp = to_long({l1 + 50, 2});
    const long bstar_exp = to_long({l1 + 59, 2});
    if (flavor == 0) {
        // mantissa as an integer-valued double, exponent lowered by 5 (src/tle.cpp:123-137)
        t.n_ddot = to_double({l1 + 44, 6}) * std::pow(10.0, static_cast<double>(nddot_exp - 5));
        t.b_star = to_double({l1 + 53, 6}) * std::pow(10.0, static_cast<double>(bstar_exp - 5));
        t.eccentricity = static_cast<double>(to_long({l2 + 26, 7})) / 1.0e7;  // tle.cpp:138
    } else {
        // implied decimal point in front of the mantissa (src/sgp4.cpp:2235-2250, 2337-2338)
        t.n_ddot = implied_decimal(l1 + 44, 5) * std::pow(10.0, static_cast<double>(nddot_exp));
        t.b_star = implied_decimal(l1 + 53, 5) * std::pow(10.0, static_cast<double>(bstar_exp));
        t.eccentricity = implied_decimal(l2 + 25, 7);
    }
    t.inclination = to_double({l2 + 8, 8});
    t.raan = to_double({l2 + 17, 8});
    t.arg_of_perigee = to_double({l2 + 34, 8});
    t.mean_anomaly = to_double({l2 + 43, 8});
    t.mean_motion = to_double({l2 + 52, 11});

    if (flavor == 0) {
        bool ok = (l1[0] == '1') && (l2[0] == '2');
        const char cls = l1[7];
        ok = ok && (cls == 'U' || cls == 'C' || cls == 'S');
        ok = ok && (1.0 <= t.epoch_day_of_year) && (t.epoch_day_of_year <= 366.0);
        ok = ok && (-15 < nddot_exp - 5) && (nddot_exp - 5 < 10);
        ok = ok && (-15 < bstar_exp - 5) && (bstar_exp - 5 < 10);
        ok = ok && (l1[62] == '0');
        ok = ok && (std::memcmp(l1 + 2, l2 + 2, 5) == 0);
        ok = ok && (0.0 <= t.inclination) && (t.inclination <= 180.0);
        ok = ok && (0.0 <= t.raan) && (t.raan <= 360.0);
        ok = ok && (0.0 <= t.arg_of_perigee) && (t.arg_of_perigee <= 360.0);
        ok = ok && (0.0 <= t.mean_anomaly) && (t.mean_anomaly <= 360.0);
        if (!ok) return 3;
        if (line_checksum(l1) != static_cast<unsigned>(l1[kLineLen - 1] - '0')) return 4;
        if (line_checksum(l2) != static_cast<unsigned>(l2[kLineLen - 1] - '0')) return 4;
    }
    return 0;
}

}  // namespace

extern "C" int perturb_gpu_parse_tles(const char *lines1, const char *lines2,
                                      size_t line_stride, size_t n, int flavor,
                                      perturb_gpu_tle *out, int32_t *status) {
    if ((n > 0 && (lines1 == nullptr || lines2 == nullptr || out == nullptr)) ||
        line_stride < static_cast<size_t>(kLineLen) || (flavor != 0 && flavor != 1)) {
        return PERTURB_GPU_ERR_ARG;
    }
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (size_t i = 0; i < n; ++i) {
        perturb_gpu_tle t;
        const int st = parse_one(lines1 + i * line_stride, lines2 + i * line_stride, flavor, t);
        // Satellite::from_tle only rejects short lines (INVALID_TLE, src/perturb.cpp:191-195);
        // TwoLineElement::parse rejects more. Either way a rejected entry becomes NaN.
        const bool reject = (flavor == 0) ? (st != 0) : (st == 2);
        if (reject) {
            t.epoch_year = t.epoch_day_of_year = t.n_dot = t.n_ddot = t.b_star = nan;
            t.inclination = t.raan = t.eccentricity = t.arg_of_perigee = nan;
            t.mean_anomaly = t.mean_motion = nan;
        }
        out[i] = t;
        if (status != nullptr) status[i] = st;
    }
    return PERTURB_GPU_OK;
}
