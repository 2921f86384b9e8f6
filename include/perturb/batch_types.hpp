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
 *   perturb::JulianDate      include/perturb/perturb.hpp:117-191 (jd, jd_frac; DateTime ctor, operator-)
 *   perturb::StateVector     include/perturb/perturb.hpp:198-205
 *   perturb::TwoLineElement  include/perturb/tle.hpp:65-91       (numeric members the path reads)
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

    /// Difference in days; the grouping preserves precision (perturb.cpp:80-83).
    double operator-(const JulianDate &rhs) const {
        return (this->jd - rhs.jd) + (this->jd_frac - rhs.jd_frac);
    }
    JulianDate operator+(const double &delta_jd) const {
        return JulianDate(jd, jd_frac + delta_jd);
    }
};

struct StateVector {
    JulianDate epoch;
    Vec3 position;  ///< TEME [km]
    Vec3 velocity;  ///< TEME [km/s]
};

/// Numeric members of a parsed TLE, TLE units (deg, rev/day).
struct TwoLineElement {
    char catalog_number[6];
    char classification;
    unsigned int epoch_year;
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
};

}  // namespace perturb

#endif  // PERTURB_PERTURB_HPP
#endif  // PERTURB_BATCH_TYPES_HPP
