/*
 * perturb/batch_types.hpp -- the public value types SatelliteBatch speaks.
 *
 * Inside a perturb source tree these come from the reference's own headers
 * (include/perturb/perturb.hpp, include/perturb/tle.hpp): define PERTURB_BATCH_USE_PERTURB_HPP
 * (or include <perturb/perturb.hpp> first) and this header adds nothing. Stand-alone -- the
 * way this repository is built and tested -- it declares layout- and name-compatible
 * equivalents so that code written against perturb compiles unchanged:
 *
 *   perturb::Sgp4Error       include/perturb/perturb.hpp:50-60   (same enumerators, values)
 *   perturb::GravModel       include/perturb/perturb.hpp:69-73
 *   perturb::Vec3            include/perturb/perturb.hpp:35
 *   perturb::DateTime        include/perturb/perturb.hpp:95-102
 *   perturb::JulianDate      include/perturb/perturb.hpp:117-191 (the whole interface: conversions
 *                            both ways, normalize, offsets, ordering)
 *   perturb::StateVector     include/perturb/perturb.hpp:198-205
 *   perturb::TwoLineElement  include/perturb/tle.hpp:65-91       (every member; parse() over the
 *                            library's exact fixed-column reader), TLEParseError, TLE_LINE_LEN
 */
#ifndef PERTURB_BATCH_TYPES_HPP
#define PERTURB_BATCH_TYPES_HPP

#if defined(PERTURB_BATCH_USE_PERTURB_HPP) && !defined(PERTURB_PERTURB_HPP)
#  include <perturb/perturb.hpp>
#endif

#ifndef PERTURB_PERTURB_HPP  // the reference's own include guard

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <string>

#include "../perturb_gpu.h"

namespace perturb {

using Vec3 = std::array<double, 3>;

enum class Sgp4Error : int {
    NONE = 0,
    MEAN_ELEMENTS,
    MEAN_MOTION,
    PERT_ELEMENTS,
    SEMI_LATUS_RECTUM,
    EPOCH_ELEMENTS_SUB_ORBITAL,
    DECAYED,
    INVALID_TLE,
    UNKNOWN
};

enum class GravModel { WGS72_OLD, WGS72, WGS84 };

struct DateTime {
    int year;    ///< 1900 .. 2100
    int month;   ///< 1 .. 12
    int day;     ///< 1 .. 31
    int hour;    ///< 0 .. 23
    int min;     ///< 0 .. 59
    double sec;  ///< 0.0 .. 59.999...
};

struct JulianDate {
    double jd;
    double jd_frac;

    JulianDate() : jd(0), jd_frac(0) {}
    /// Calendar date to (whole day at midnight, fraction of day): the arithmetic of
    /// sgp4::jday_SGP4 (sgp4.cpp:3100-3121), which the reference's constructor calls
    /// (perturb.cpp:47-52), operation for operation.
    explicit JulianDate(DateTime t) {
        jd = 367.0 * t.year - std::floor((7 * (t.year + std::floor((t.month + 9) / 12.0))) * 0.25) +
             std::floor(275 * t.month / 9.0) + t.day + 1721013.5;
        jd_frac = (t.sec + t.min * 60.0 + t.hour * 3600.0) / 86400.0;
        if (std::fabs(jd_frac) > 1.0) {
            const double whole = std::floor(jd_frac);
            jd = jd + whole;
            jd_frac = jd_frac - whole;
        }
    }
    explicit JulianDate(double _jd) : jd(_jd), jd_frac(0) {}
    explicit JulianDate(double _jd, double _jd_frac) : jd(_jd), jd_frac(_jd_frac) {}

    /// Same time point as a calendar date (perturb.cpp:54-58): the arithmetic of
    /// sgp4::invjday_SGP4 and sgp4::days2mdhms_SGP4 (sgp4.cpp:3232-3279, 3161-3192).
    DateTime to_datetime() const {
        double day_no = jd, part = jd_frac;
        if (std::fabs(part) >= 1.0) {  // whole days hiding in the fraction
            day_no = day_no + std::floor(part);
            part = part - std::floor(part);
        }
        const double off = day_no - std::floor(day_no) - 0.5;  // fraction hiding in the day
        if (std::fabs(off) > 0.00000001) {
            day_no = day_no - off;
            part = part + off;
        }
        const double since1900 = day_no - 2415019.5;
        DateTime t = DateTime();
        t.year = 1900 + static_cast<int>(std::floor(since1900 / 365.25));
        int leaps = static_cast<int>(std::floor((t.year - 1901) * 0.25));
        double days = std::floor(since1900 - ((t.year - 1900) * 365.0 + leaps));
        if (days + part < 1.0) {  // the very beginning of a year belongs to the year before
            t.year = t.year - 1;
            leaps = static_cast<int>(std::floor((t.year - 1901) * 0.25));
            days = std::floor(since1900 - ((t.year - 1900) * 365.0 + leaps));
        }
        // day of year (with fraction) -> month, day, hour, minute, second
        const double doy = days + part;
        const int whole = static_cast<int>(std::floor(doy));
        const int feb = (t.year % 4) == 0 ? 29 : 28;
        const int len[12] = {31, feb, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
        int m = 0, before = 0;
        while (whole > before + len[m] && m < 11) {
            before += len[m];
            ++m;
        }
        t.month = m + 1;
        t.day = whole - before;
        double rest = (doy - whole) * 24.0;
        t.hour = static_cast<int>(std::floor(rest));
        rest = (rest - t.hour) * 60.0;
        t.min = static_cast<int>(std::floor(rest));
        t.sec = (rest - t.min) * 60.0;
        return t;
    }

    /// Whole days in `jd` (at midnight, x.5), a [0, 1) offset in `jd_frac` (perturb.cpp:60-73).
    void normalize() {
        // a fraction of a day hiding in `jd` (anything off the x.5 midnight grid by more than
        // 1e-12) moves into `jd_frac` ...
        const double off_grid = jd - std::floor(jd) - 0.5;
        if (std::fabs(off_grid) > 1e-12) {
            jd = jd - off_grid;
            jd_frac = jd_frac + off_grid;
        }
        // ... and whole days hiding in `jd_frac` move into `jd`
        if (std::fabs(jd_frac) >= 1.0) {
            const double days = std::floor(jd_frac);
            jd = jd + days;
            jd_frac = jd_frac - days;
        }
    }
    JulianDate normalized() const {
        JulianDate other = *this;
        other.normalize();
        return other;
    }

    /// Difference in days; the grouping preserves precision (perturb.cpp:80-83).
    double operator-(const JulianDate &rhs) const {
        return (this->jd - rhs.jd) + (this->jd_frac - rhs.jd_frac);
    }
    /// Offset by days, added to `jd_frac` only: not normalized (perturb.cpp:85-101).
    JulianDate operator+(const double &delta_jd) const {
        return JulianDate(jd, jd_frac + delta_jd);
    }
    JulianDate &operator+=(const double &delta_jd) { return *this = *this + delta_jd; }
    JulianDate operator-(const double &delta_jd) const { return *this + (-delta_jd); }
    JulianDate &operator-=(const double &delta_jd) { return *this = *this - delta_jd; }

    /// Chronological order through the precise difference (perturb.cpp:103-117).
    bool operator<(const JulianDate &rhs) const { return (*this - rhs) < 0; }
    bool operator>(const JulianDate &rhs) const { return (*this - rhs) > 0; }
    bool operator<=(const JulianDate &rhs) const { return (*this - rhs) <= 0; }
    bool operator>=(const JulianDate &rhs) const { return (*this - rhs) >= 0; }
};

struct StateVector {
    JulianDate epoch;
    Vec3 position;  ///< TEME [km]
    Vec3 velocity;  ///< TEME [km/s]
};

/// Classical Keplerian elements, member for member the reference's
/// (include/perturb/perturb.hpp:211-222; [km] and [rad]). The batch computes them for whole
/// propagate results on the device (`SatelliteBatch::classical_elements_soa`, eleven planes
/// in this member order); there is no per-state host constructor here -- no CPU path.
struct ClassicalOrbitalElements {
    double semilatus_rectum;
    double semimajor_axis;
    double eccentricity;
    double inclination;
    double raan;
    double arg_of_perigee;
    double true_anomaly;
    double mean_anomaly;
    double arg_of_latitude;
    double true_longitude;
    double longitude_of_periapsis;
};

/// Number of characters of one TLE line (include/perturb/tle.hpp:33).
constexpr std::size_t TLE_LINE_LEN = 69;

/// Outcome of `TwoLineElement::parse`, in checking order (include/perturb/tle.hpp:45-51).
enum class TLEParseError {
    NONE,
    SHOULD_BE_SPACE,
    INVALID_FORMAT,
    INVALID_VALUE,
    CHECKSUM_MISMATCH,
};

/// A parsed TLE record, TLE units (deg, rev/day): member for member the reference's
/// (include/perturb/tle.hpp:65-91). The batch reads the numeric ones.
struct TwoLineElement {
    // Line 1
    char catalog_number[6];
    char classification;
    unsigned int launch_year;
    unsigned int launch_number;
    char launch_piece[4];
    unsigned int epoch_year;
    double epoch_day_of_year;
    double n_dot;
    double n_ddot;
    double b_star;
    unsigned char ephemeris_type;
    unsigned int element_set_number;
    unsigned char line_1_checksum;
    // Line 2
    double inclination;
    double raan;
    double eccentricity;
    double arg_of_perigee;
    double mean_anomaly;
    double mean_motion;
    unsigned long revolution_number;
    unsigned char line_2_checksum;

    /// `TwoLineElement::parse` (src/tle.cpp:41-173) over the library's restatement of the
    /// reference's sscanf directives (`perturb_gpu_parse_tle`): same values bit for bit, same
    /// error in the same checking order, including the lines the reference rejects because a
    /// blank field shifts its tokens. Lines must hold TLE_LINE_LEN characters.
    TLEParseError parse(const char *line_1, const char *line_2) {
        perturb_gpu_tle num;
        perturb_gpu_tle_ident id;
        const int status = perturb_gpu_parse_tle(line_1, line_2, &num, &id);
        if (status < 0) {
            return TLEParseError::INVALID_FORMAT;
        }
        if (status == 0 || status >= 3) {  // every field was read
            for (int i = 0; i < 6; ++i) {
                catalog_number[i] = id.catalog_number[i];
            }
            classification = id.classification;
            launch_year = id.launch_year;
            launch_number = id.launch_number;
            for (int i = 0; i < 4; ++i) {
                launch_piece[i] = id.launch_piece[i];
            }
            ephemeris_type = id.ephemeris_type;
            element_set_number = id.element_set_number;
            line_1_checksum = id.line_1_checksum;
            revolution_number = id.revolution_number;
            line_2_checksum = id.line_2_checksum;
            epoch_year = static_cast<unsigned int>(num.epoch_year);
            epoch_day_of_year = num.epoch_day_of_year;
            n_dot = num.n_dot;
            n_ddot = num.n_ddot;
            b_star = num.b_star;
            inclination = num.inclination;
            raan = num.raan;
            eccentricity = num.eccentricity;
            arg_of_perigee = num.arg_of_perigee;
            mean_anomaly = num.mean_anomaly;
            mean_motion = num.mean_motion;
        }
        return static_cast<TLEParseError>(status);
    }
    TLEParseError parse(const std::string &line_1, const std::string &line_2) {
        if (line_1.length() < TLE_LINE_LEN || line_2.length() < TLE_LINE_LEN) {
            return TLEParseError::INVALID_FORMAT;
        }
        return this->parse(line_1.c_str(), line_2.c_str());
    }
};

}  // namespace perturb

#endif  // PERTURB_PERTURB_HPP
#endif  // PERTURB_BATCH_TYPES_HPP
