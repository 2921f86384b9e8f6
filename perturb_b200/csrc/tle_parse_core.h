// Fixed-column TLE reader shared by the host entry point (perturb_gpu_parse_tles) and the
// device kernel (perturb_gpu_init_from_text): ONE implementation, so host and device parse to
// the same bits by construction.
//
// The reference has two readers: perturb's TwoLineElement::parse (src/tle.cpp:41-173, sscanf
// based) and Vallado's twoline2rv (src/sgp4.cpp:2231-2338, in-place string fix-ups + sscanf).
// A TLE is a fixed-column card image, so columns are sliced directly. Decimal -> double uses
// the exact fast path of strtod: a field has at most 12 significant digits, so its digits M
// (< 2^53) and 10^f (f <= 22) are both exact doubles and ONE IEEE division M / 10^f is the
// correctly rounded value -- bit-identical to sscanf("%lf") / strtod on the same characters.
// pow(10.0, exp) factors come from a table the host fills with std::pow (what both reference
// readers call), so the scaled n_ddot / b_star match the reference bit for bit as well.
#ifndef PERTURB_B200_TLE_PARSE_CORE_H
#define PERTURB_B200_TLE_PARSE_CORE_H

#include <stdint.h>

#include "../../include/perturb_gpu.h"

#if defined(__CUDACC__)
#  define PB_HD __host__ __device__
#else
#  define PB_HD
#endif

namespace pb {
namespace tle {

constexpr int kLineLen = 69;      // perturb::TLE_LINE_LEN, include/perturb/tle.hpp:33
constexpr int kPowBias = 30;      // table index = exponent + kPowBias
constexpr int kPowCount = 61;     // exponents -30 .. +30

struct Pow10Table {
    double p[kPowCount];  // p[e + kPowBias] = std::pow(10.0, e), filled on the host
};

PB_HD inline double exact_pow10(int f) {  // 10^f, exactly representable for 0 <= f <= 22
    const double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                          1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[f < 0 ? 0 : (f > 22 ? 22 : f)];
}

PB_HD inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// strtod on the characters [p, p+len): leading blanks, optional sign, digits with at most
// one '.', stopping at the first character that does not fit. No digits -> 0.0.
PB_HD inline double slice_to_double(const char *p, int len) {
    int i = 0;
    while (i < len && is_space(p[i])) ++i;
    bool neg = false;
    if (i < len && (p[i] == '+' || p[i] == '-')) {
        neg = (p[i] == '-');
        ++i;
    }
    uint64_t mant = 0;
    int frac = 0, digits = 0;
    bool seen_point = false;
    for (; i < len; ++i) {
        const char c = p[i];
        if (c >= '0' && c <= '9') {
            if (digits < 18) {
                mant = mant * 10u + static_cast<uint64_t>(c - '0');
                if (seen_point) ++frac;
                if (mant != 0) ++digits;
            } else if (!seen_point) {
                --frac;  // cannot happen in a 12-column field; keeps the magnitude sane
            }
        } else if (c == '.' && !seen_point) {
            seen_point = true;
        } else {
            break;
        }
    }
    double v = static_cast<double>(mant);
    if (frac > 0) {
        v = v / exact_pow10(frac);
    } else if (frac < 0) {
        v = v * exact_pow10(-frac);
    }
    return neg ? -v : v;
}

// strtol(.., 10) on the slice.
PB_HD inline long slice_to_long(const char *p, int len) {
    int i = 0;
    while (i < len && is_space(p[i])) ++i;
    bool neg = false;
    if (i < len && (p[i] == '+' || p[i] == '-')) {
        neg = (p[i] == '-');
        ++i;
    }
    long v = 0;
    for (; i < len && p[i] >= '0' && p[i] <= '9'; ++i) v = v * 10 + (p[i] - '0');
    return neg ? -v : v;
}

// "[sign]ddddd" with the decimal point implied in front of the digits (blank digit = 0, as
// twoline2rv's fix-ups make it, src/sgp4.cpp:2242-2253): value [-]0.ddddd
PB_HD inline double implied_decimal(const char *sign_col, int digits) {
    uint64_t mant = 0;
    for (int i = 1; i <= digits; ++i) {
        const char c = sign_col[i];
        if (c >= '0' && c <= '9') {
            mant = mant * 10u + static_cast<uint64_t>(c - '0');
        } else if (c == ' ') {
            mant = mant * 10u;
        } else {  // strtod would stop here: the remaining places count as not present
            double v = static_cast<double>(mant) / exact_pow10(i - 1);
            return (*sign_col == '-') ? -v : v;
        }
    }
    const double v = static_cast<double>(mant) / exact_pow10(digits);
    return (*sign_col == '-') ? -v : v;
}

PB_HD inline double pow10_of(const Pow10Table &pw, long e) {
    if (e < -kPowBias) e = -kPowBias;
    if (e > kPowBias) e = kPowBias;
    return pw.p[e + kPowBias];
}

PB_HD inline unsigned line_checksum(const char *line) {  // src/tle.cpp:25-36
    unsigned sum = 0U;
    for (int i = 0; i < kLineLen - 1; ++i) {
        const char c = line[i];
        if (c >= '0' && c <= '9') sum += static_cast<unsigned>(c - '0');
        if (c == '-') sum += 1U;
    }
    return sum % 10U;
}

PB_HD inline bool has_nul(const char *line) {
    for (int i = 0; i < kLineLen; ++i) {
        if (line[i] == '\0') return true;
    }
    return false;
}

// Returns a perturb::TLEParseError value (0 NONE, 1 SHOULD_BE_SPACE, 2 INVALID_FORMAT,
// 3 INVALID_VALUE, 4 CHECKSUM_MISMATCH), checked in that order like src/tle.cpp:39-44 says.
// flavor 0: TwoLineElement::parse's numeric conventions; flavor 1: twoline2rv's.
PB_HD inline int parse_one(const char *l1, const char *l2, int flavor, const Pow10Table &pw,
                           perturb_gpu_tle &t) {
    if (has_nul(l1) || has_nul(l2)) return 2;  // shorter than TLE_LINE_LEN
    if (flavor == 0) {
        const int spaces1[8] = {2, 9, 18, 33, 44, 53, 62, 64};  // src/tle.cpp:43
        const int spaces2[7] = {2, 8, 17, 26, 34, 43, 52};      // src/tle.cpp:49
        for (int i = 0; i < 8; ++i) {
            if (l1[spaces1[i] - 1] != ' ') return 1;
        }
        for (int i = 0; i < 7; ++i) {
            if (l2[spaces2[i] - 1] != ' ') return 1;
        }
    }

    t.epoch_year = static_cast<double>(slice_to_long(l1 + 18, 2));
    t.epoch_day_of_year = slice_to_double(l1 + 20, 12);
    t.n_dot = slice_to_double(l1 + 33, 10);
    const long nddot_exp = slice_to_long(l1 + 50, 2);
    const long bstar_exp = slice_to_long(l1 + 59, 2);
    if (flavor == 0) {
        // mantissa as an integer-valued double, exponent lowered by 5 (src/tle.cpp:123-137)
        t.n_ddot = slice_to_double(l1 + 44, 6) * pow10_of(pw, nddot_exp - 5);
        t.b_star = slice_to_double(l1 + 53, 6) * pow10_of(pw, bstar_exp - 5);
        t.eccentricity = static_cast<double>(slice_to_long(l2 + 26, 7)) / 1.0e7;  // tle.cpp:138
    } else {
        // implied decimal point in front of the mantissa (src/sgp4.cpp:2235-2250, 2337-2338)
        t.n_ddot = implied_decimal(l1 + 44, 5) * pow10_of(pw, nddot_exp);
        t.b_star = implied_decimal(l1 + 53, 5) * pow10_of(pw, bstar_exp);
        t.eccentricity = implied_decimal(l2 + 25, 7);
    }
    t.inclination = slice_to_double(l2 + 8, 8);
    t.raan = slice_to_double(l2 + 17, 8);
    t.arg_of_perigee = slice_to_double(l2 + 34, 8);
    t.mean_anomaly = slice_to_double(l2 + 43, 8);
    t.mean_motion = slice_to_double(l2 + 52, 11);

    if (flavor == 0) {
        bool ok = (l1[0] == '1') && (l2[0] == '2');
        const char cls = l1[7];
        ok = ok && (cls == 'U' || cls == 'C' || cls == 'S');
        ok = ok && (1.0 <= t.epoch_day_of_year) && (t.epoch_day_of_year <= 366.0);
        ok = ok && (-15 < nddot_exp - 5) && (nddot_exp - 5 < 10);
        ok = ok && (-15 < bstar_exp - 5) && (bstar_exp - 5 < 10);
        ok = ok && (l1[62] == '0');
        for (int i = 2; i < 7; ++i) ok = ok && (l1[i] == l2[i]);
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

// Satellite::from_tle only rejects short lines (INVALID_TLE, src/perturb.cpp:191-195);
// TwoLineElement::parse rejects more. A rejected entry becomes all-NaN, which kernel 1 turns
// into a zeroed record with error INVALID_TLE.
PB_HD inline void finish_entry(int status, int flavor, perturb_gpu_tle &t) {
    const bool reject = (flavor == 0) ? (status != 0) : (status == 2);
    if (reject) {
#if defined(__CUDA_ARCH__)
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
#else
        const double nan = __builtin_nan("");
#endif
        t.epoch_year = t.epoch_day_of_year = t.n_dot = t.n_ddot = t.b_star = nan;
        t.inclination = t.raan = t.eccentricity = t.arg_of_perigee = nan;
        t.mean_anomaly = t.mean_motion = nan;
    }
}

}  // namespace tle
}  // namespace pb

#endif
