// Fixed-column TLE reader shared by the host entry point (perturb_gpu_parse_tles) and the
// device kernel (perturb_gpu_init_from_text): ONE implementation, so host and device parse to
// the same bits by construction.
//
// The reference has two readers: perturb's TwoLineElement::parse (src/tle.cpp:41-173, sscanf
// based) and Vallado's twoline2rv (src/sgp4.cpp:2231-2338, in-place string fix-ups + sscanf).
// A TLE is a fixed-column card image, so twoline2rv's columns are sliced directly; perturb's
// reader lets blank fields shift its tokens, so its directives are restated (struct Scan). Decimal -> double uses
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

// ---- TwoLineElement::parse (flavor 0) reads its lines with sscanf (src/tle.cpp:57-99) ----
// The directives are restated here over a 69-character window so that a field that is blank
// or out of place shifts the tokens exactly as it does in the reference (which then reports
// INVALID_FORMAT / INVALID_VALUE for lines a column reader would accept, e.g. a blank
// international designator). Supported conversions: %Nu / %Nd / %Nlu (optional sign counted
// in the width), %Ns, %1c, %Nlf (decimal digits with at most one point; the correctly rounded
// value, like strtod), white space, %n. Not restated: exponents, "inf", "nan" and hexadecimal
// floats inside a %lf -- none can occur in a TLE that passes the later checks.
struct Scan {
    const char *s;
    int n, pos;
    PB_HD void skip() {
        while (pos < n && is_space(s[pos])) ++pos;
    }
    // %Nd / %Nu: false on a matching or input failure
    PB_HD bool integer(int width, long &v) {
        skip();
        if (pos >= n) return false;
        int used = 0;
        bool neg = false;
        if (s[pos] == '+' || s[pos] == '-') {
            neg = (s[pos] == '-');
            ++pos;
            ++used;
        }
        long acc = 0;
        int digits = 0;
        while (pos < n && used < width && s[pos] >= '0' && s[pos] <= '9') {
            acc = acc * 10 + (s[pos] - '0');
            ++pos;
            ++used;
            ++digits;
        }
        if (digits == 0) return false;
        v = neg ? -acc : acc;
        return true;
    }
    PB_HD bool word(int width, char *out) {  // %Ns
        skip();
        if (pos >= n) return false;
        int k = 0;
        while (pos < n && k < width && !is_space(s[pos])) out[k++] = s[pos++];
        out[k] = '\0';
        return true;
    }
    PB_HD bool chr(char &c) {  // %1c
        if (pos >= n) return false;
        c = s[pos++];
        return true;
    }
    PB_HD bool real(int width, double &v) {  // %Nlf
        skip();
        if (pos >= n) return false;
        const int start = pos;
        int used = 0, digits = 0;
        bool point = false;
        if (s[pos] == '+' || s[pos] == '-') {
            ++pos;
            ++used;
        }
        while (pos < n && used < width) {
            const char c = s[pos];
            if (c >= '0' && c <= '9') {
                ++digits;
            } else if (c == '.' && !point) {
                point = true;
            } else {
                break;
            }
            ++pos;
            ++used;
        }
        if (digits == 0) return false;
        v = slice_to_double(s + start, pos - start);
        return true;
    }
};

// Returns a perturb::TLEParseError value (0 NONE, 1 SHOULD_BE_SPACE, 2 INVALID_FORMAT,
// 3 INVALID_VALUE, 4 CHECKSUM_MISMATCH), checked in that order like src/tle.cpp:39-44 says.
// flavor 0: TwoLineElement::parse's numeric conventions; flavor 1: twoline2rv's.
PB_HD inline int parse_perturb(const char *l1, const char *l2, const Pow10Table &pw,
                               perturb_gpu_tle &t, perturb_gpu_tle_ident *ident = nullptr) {
    const int spaces1[8] = {2, 9, 18, 33, 44, 53, 62, 64};  // src/tle.cpp:43
    const int spaces2[7] = {2, 8, 17, 26, 34, 43, 52};      // src/tle.cpp:49
    for (int i = 0; i < 8; ++i) {
        if (l1[spaces1[i] - 1] != ' ') return 1;
    }
    for (int i = 0; i < 7; ++i) {
        if (l2[spaces2[i] - 1] != ' ') return 1;
    }
    // "%1hhu %5s %1c %2u %3u %3s %2u %12lf %10lf %6lf %2d %6lf %2d %1hhu %4u %n %1hhu"
    Scan a = {l1, kLineLen, 0};
    long line1_num = 0, launch_year = 0, launch_number = 0, epoch_year = 0, nddot_exp = 0;
    long bstar_exp = 0, ephemeris = 0, elset = 0, sum1 = 0;
    char cat1[6] = {0}, piece[4] = {0}, cls = 0;
    double epoch_day = 0.0, n_dot = 0.0, n_ddot = 0.0, b_star = 0.0;
    int got1 = 0, pre1 = 0;
    do {
        if (!a.integer(1, line1_num)) break;
        ++got1;
        if (!a.word(5, cat1)) break;
        ++got1;
        a.skip();
        if (!a.chr(cls)) break;
        ++got1;
        if (!a.integer(2, launch_year)) break;
        ++got1;
        if (!a.integer(3, launch_number)) break;
        ++got1;
        if (!a.word(3, piece)) break;
        ++got1;
        if (!a.integer(2, epoch_year)) break;
        ++got1;
        if (!a.real(12, epoch_day)) break;
        ++got1;
        if (!a.real(10, n_dot)) break;
        ++got1;
        if (!a.real(6, n_ddot)) break;
        ++got1;
        if (!a.integer(2, nddot_exp)) break;
        ++got1;
        if (!a.real(6, b_star)) break;
        ++got1;
        if (!a.integer(2, bstar_exp)) break;
        ++got1;
        if (!a.integer(1, ephemeris)) break;
        ++got1;
        if (!a.integer(4, elset)) break;
        ++got1;
        a.skip();
        pre1 = a.pos;
        if (!a.integer(1, sum1)) break;
        ++got1;
    } while (false);
    // "%1hhu %5s %8lf %8lf %7lu %8lf %8lf %11lf %5lu %n %1hhu" (%10lf when column 53 is blank)
    Scan b = {l2, kLineLen, 0};
    long line2_num = 0, ecc_int = 0, revs = 0, sum2 = 0;
    char cat2[6] = {0};
    double incl = 0.0, raan = 0.0, argp = 0.0, manom = 0.0, mmotion = 0.0;
    int got2 = 0, pre2 = 0;
    do {
        if (!b.integer(1, line2_num)) break;
        ++got2;
        if (!b.word(5, cat2)) break;
        ++got2;
        if (!b.real(8, incl)) break;
        ++got2;
        if (!b.real(8, raan)) break;
        ++got2;
        if (!b.integer(7, ecc_int)) break;
        ++got2;
        if (!b.real(8, argp)) break;
        ++got2;
        if (!b.real(8, manom)) break;
        ++got2;
        if (!b.real(l2[52] != ' ' ? 11 : 10, mmotion)) break;
        ++got2;
        if (!b.integer(5, revs)) break;
        ++got2;
        b.skip();
        pre2 = b.pos;
        if (!b.integer(1, sum2)) break;
        ++got2;
    } while (false);
    // an element set / revolution number without leading zeros swallows the checksum digit
    // (src/tle.cpp:105-127)
    if (got1 == 15 && pre1 >= kLineLen && l1[kLineLen - 5] == ' ') {
        elset = static_cast<long>(static_cast<unsigned int>(elset) / 10U);  // an unsigned member
        sum1 = l1[kLineLen - 1] - '0';
        ++got1;
    }
    if (got2 == 9 && pre2 >= kLineLen && l2[kLineLen - 6] == ' ') {
        revs = static_cast<long>(static_cast<unsigned long>(revs) / 10UL);
        sum2 = l2[kLineLen - 1] - '0';
        ++got2;
    }
    if (l1[44] != '.') nddot_exp -= 5;
    if (l1[53] != '.') bstar_exp -= 5;
    if (got1 != 16 || got2 != 10) return 2;

    t.epoch_year = static_cast<double>(static_cast<unsigned int>(epoch_year));
    t.epoch_day_of_year = epoch_day;
    t.n_dot = n_dot;
    t.n_ddot = n_ddot * pow10_of(pw, nddot_exp);
    t.b_star = b_star * pow10_of(pw, bstar_exp);
    t.eccentricity = static_cast<double>(static_cast<unsigned long>(ecc_int)) / 1.0e7;
    t.inclination = incl;
    t.raan = raan;
    t.arg_of_perigee = argp;
    t.mean_anomaly = manom;
    t.mean_motion = mmotion;

    // src/tle.cpp:140-162 (unsigned members: a scanned minus sign wraps to a huge value)
    bool ok = (static_cast<unsigned char>(line1_num) == 1);
    ok = ok && (cls == 'U' || cls == 'C' || cls == 'S');
    ok = ok && (static_cast<unsigned int>(launch_year) < 100U) &&
         (static_cast<unsigned int>(epoch_year) < 100U);
    ok = ok && (1.0 <= epoch_day) && (epoch_day <= 366.0);
    ok = ok && (-15 < nddot_exp) && (nddot_exp < 10);
    ok = ok && (-15 < bstar_exp) && (bstar_exp < 10);
    ok = ok && (static_cast<unsigned char>(ephemeris) == 0U);
    ok = ok && (static_cast<unsigned int>(elset) < 10000U);
    ok = ok && (static_cast<unsigned char>(line2_num) == 2);
    for (int i = 0; i < 6; ++i) {
        ok = ok && (cat1[i] == cat2[i]);
        if (cat1[i] == '\0' || cat2[i] == '\0') break;
    }
    ok = ok && (0.0 <= incl) && (incl <= 180.0);
    ok = ok && (0.0 <= raan) && (raan <= 360.0);
    ok = ok && (0.0 <= argp) && (argp <= 360.0);
    ok = ok && (0.0 <= manom) && (manom <= 360.0);
    if (ident != nullptr) {  // the members of perturb::TwoLineElement that are not numbers
        for (int i = 0; i < 6; ++i) ident->catalog_number[i] = cat1[i];
        ident->catalog_number[5] = '\0';
        ident->classification = cls;
        ident->launch_year = static_cast<unsigned int>(launch_year);
        ident->launch_number = static_cast<unsigned int>(launch_number);
        for (int i = 0; i < 4; ++i) ident->launch_piece[i] = piece[i];
        ident->launch_piece[3] = '\0';
        ident->ephemeris_type = static_cast<unsigned char>(ephemeris);
        ident->element_set_number = static_cast<unsigned int>(elset);
        ident->line_1_checksum = static_cast<unsigned char>(sum1);
        ident->revolution_number = static_cast<unsigned long>(revs);
        ident->line_2_checksum = static_cast<unsigned char>(sum2);
    }
    if (!ok) return 3;
    if (line_checksum(l1) != static_cast<unsigned>(static_cast<unsigned char>(sum1))) return 4;
    if (line_checksum(l2) != static_cast<unsigned>(static_cast<unsigned char>(sum2))) return 4;
    return 0;
}

PB_HD inline int parse_one(const char *l1, const char *l2, int flavor, const Pow10Table &pw,
                           perturb_gpu_tle &t) {
    if (has_nul(l1) || has_nul(l2)) return 2;  // shorter than TLE_LINE_LEN
    if (flavor == 0) return parse_perturb(l1, l2, pw, t);

    t.epoch_year = static_cast<double>(slice_to_long(l1 + 18, 2));
    t.epoch_day_of_year = slice_to_double(l1 + 20, 12);
    t.n_dot = slice_to_double(l1 + 33, 10);
    const long nddot_exp = slice_to_long(l1 + 50, 2);
    const long bstar_exp = slice_to_long(l1 + 59, 2);
    // implied decimal point in front of the mantissa (src/sgp4.cpp:2235-2250, 2337-2338)
    t.n_ddot = implied_decimal(l1 + 44, 5) * pow10_of(pw, nddot_exp);
    t.b_star = implied_decimal(l1 + 53, 5) * pow10_of(pw, bstar_exp);
    t.eccentricity = implied_decimal(l2 + 25, 7);
    t.inclination = slice_to_double(l2 + 8, 8);
    t.raan = slice_to_double(l2 + 17, 8);
    t.arg_of_perigee = slice_to_double(l2 + 34, 8);
    t.mean_anomaly = slice_to_double(l2 + 43, 8);
    t.mean_motion = slice_to_double(l2 + 52, 11);
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
