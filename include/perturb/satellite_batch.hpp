/*
 * perturb/satellite_batch.hpp -- perturb::SatelliteBatch, the batched counterpart of
 * perturb::Satellite (include/perturb/perturb.hpp:241-322) running on B200 GPUs.
 *
 * A loop such as
 *
 *     std::vector<Satellite> sats = ...;                 // Satellite(tle, grav) each
 *     for (auto &sat : sats)
 *         for (auto &t : times) errs[..] = sat.propagate(t, svs[..]);   // perturb.cpp:236-242
 *
 * becomes
 *
 *     SatelliteBatch batch(tles.data(), tles.size(), GravModel::WGS72);
 *     batch.propagate(times.data(), times.size(), svs.data(), errs.data());
 *
 * C++11, header-only over the C ABI (include/perturb_gpu.h), no exceptions, no RTTI; the
 * caller owns every buffer. Construction problems are reported like the reference does: check
 * ok() / status() for library failures and last_error(i) for per-satellite Sgp4Error codes.
 */
#ifndef PERTURB_SATELLITE_BATCH_HPP
#define PERTURB_SATELLITE_BATCH_HPP

#include <cstddef>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "../perturb_gpu.h"
#include "batch_types.hpp"

namespace perturb {

class SatelliteBatch {
public:
    /// Batched `Satellite(const TwoLineElement &, GravModel)` (perturb.cpp:135-186).
    SatelliteBatch(
        const TwoLineElement *tles, std::size_t n, GravModel grav_model = GravModel::WGS72,
        int device = 0
    )
        : handle_(nullptr), n_(n), status_(PERTURB_GPU_OK) {
        std::vector<perturb_gpu_tle> packed(n);
        for (std::size_t i = 0; i < n; ++i) {
            packed[i] = pack(tles[i]);
        }
        init(packed, grav_model, device);
    }

    /// Batched `Satellite::from_tle(line_1, line_2, grav)` (perturb.cpp:189-217): lines are
    /// read with Vallado's numeric conventions; a line shorter than 69 characters gives that
    /// satellite `Sgp4Error::INVALID_TLE`, as in the reference.
    static SatelliteBatch from_tles(
        const std::string *lines_1, const std::string *lines_2, std::size_t n,
        GravModel grav_model = GravModel::WGS72, int device = 0
    ) {
        constexpr std::size_t STRIDE = 70;
        std::vector<char> b1(n * STRIDE, '\0'), b2(n * STRIDE, '\0');
        for (std::size_t i = 0; i < n; ++i) {
            lines_1[i].copy(&b1[i * STRIDE], STRIDE - 1);
            lines_2[i].copy(&b2[i * STRIDE], STRIDE - 1);
        }
        SatelliteBatch batch;
        batch.n_ = n;
        batch.init_error_.assign(n, static_cast<std::int32_t>(Sgp4Error::UNKNOWN));
        // TLE text is parsed on the device (flavor 1 = twoline2rv's numeric conventions),
        // 'i' = the opsmode Satellite::from_tle hard-wires (perturb.cpp:200)
        batch.status_ = perturb_gpu_init_from_text(
            b1.data(), b2.data(), STRIDE, n, 1, grav_code(grav_model), 'i', device,
            &batch.handle_, batch.init_error_.data(), nullptr
        );
        if (batch.status_ == PERTURB_GPU_OK) {
            batch.fetch_queries();
        }
        return batch;
    }

    ~SatelliteBatch() { perturb_gpu_free(handle_); }

    SatelliteBatch(SatelliteBatch &&other) noexcept
        : handle_(other.handle_), n_(other.n_), status_(other.status_),
          init_error_(std::move(other.init_error_)), epoch_jd_(std::move(other.epoch_jd_)),
          epoch_frac_(std::move(other.epoch_frac_)) {
        other.handle_ = nullptr;
    }
    SatelliteBatch &operator=(SatelliteBatch &&other) noexcept {
        if (this != &other) {
            perturb_gpu_free(handle_);
            handle_ = other.handle_;
            n_ = other.n_;
            status_ = other.status_;
            init_error_ = std::move(other.init_error_);
            epoch_jd_ = std::move(other.epoch_jd_);
            epoch_frac_ = std::move(other.epoch_frac_);
            other.handle_ = nullptr;
        }
        return *this;
    }
    SatelliteBatch(const SatelliteBatch &) = delete;
    SatelliteBatch &operator=(const SatelliteBatch &) = delete;

    /// False if the library itself failed (no GPU, allocation, bad arguments).
    bool ok() const { return status_ == PERTURB_GPU_OK && (handle_ != nullptr || n_ == 0); }
    int status() const { return status_; }
    static const char *status_message() { return perturb_gpu_strerror(); }

    std::size_t size() const { return n_; }

    /// `Satellite::last_error()` right after construction (perturb.cpp:220-222).
    Sgp4Error last_error(std::size_t i) const {
        return i < init_error_.size() ? static_cast<Sgp4Error>(init_error_[i])
                                      : Sgp4Error::UNKNOWN;
    }

    /// `Satellite::epoch()` (perturb.cpp:224-226).
    JulianDate epoch(std::size_t i) const {
        return i < epoch_jd_.size() ? JulianDate(epoch_jd_[i], epoch_frac_[i]) : JulianDate();
    }

    /// Batched `Satellite::propagate(jd, sv)` for every satellite and every time.
    /// `svs` and `errs` are satellite-major [size() * m]; a StateVector is left untouched
    /// for error codes 1-4, exactly like the reference (sgp4.cpp:1865,1878,1943,1995).
    /// For throughput-critical callers use propagate_soa(), which avoids this transposition.
    int propagate(const JulianDate *times, std::size_t m, StateVector *svs, Sgp4Error *errs) {
        std::vector<double> jd(m), frac(m);
        for (std::size_t k = 0; k < m; ++k) {
            jd[k] = times[k].jd;
            frac[k] = times[k].jd_frac;
        }
        std::vector<double> posvel(6 * n_ * m);
        std::vector<std::int32_t> codes(n_ * m);
        const int st = propagate_soa(
            jd.data(), frac.data(), m, posvel.data(), codes.data(), m, n_ * m, nullptr
        );
        if (st != PERTURB_GPU_OK) {
            return st;
        }
        scatter(posvel, codes, m, svs, errs);
        for (std::size_t i = 0; i < n_; ++i) {
            for (std::size_t k = 0; k < m; ++k) {
                svs[i * m + k].epoch = times[k];  // perturb.cpp:240
            }
        }
        return st;
    }

    /// Batched `Satellite::propagate_from_epoch(mins, sv)` (perturb.cpp:228-234).
    int propagate_from_epoch(
        const double *mins, std::size_t m, StateVector *svs, Sgp4Error *errs
    ) {
        std::vector<double> posvel(6 * n_ * m);
        std::vector<std::int32_t> codes(n_ * m);
        status_ = perturb_gpu_propagate_from_epoch(
            handle_, mins, m, posvel.data(), codes.data(), m, n_ * m, nullptr
        );
        if (status_ == PERTURB_GPU_OK) {
            status_ = perturb_gpu_synchronize(handle_);
        }
        if (status_ != PERTURB_GPU_OK) {
            return status_;
        }
        scatter(posvel, codes, m, svs, errs);
        for (std::size_t i = 0; i < n_; ++i) {
            for (std::size_t k = 0; k < m; ++k) {
                svs[i * m + k].epoch = epoch(i) + (mins[k] / 1440.0);  // perturb.cpp:229
            }
        }
        return status_;
    }

    /// Structure-of-arrays variant, host or device buffers (see perturb_gpu_propagate):
    /// posvel[c * plane_stride + i * row_stride + k], err[i * row_stride + k].
    int propagate_soa(
        const double *jd, const double *jd_frac, std::size_t m, double *posvel,
        std::int32_t *err, std::size_t row_stride, std::size_t plane_stride, void *stream
    ) {
        status_ = perturb_gpu_propagate(
            handle_, jd, jd_frac, m, posvel, err, row_stride, plane_stride, stream
        );
        if (status_ == PERTURB_GPU_OK && stream == nullptr) {
            status_ = perturb_gpu_synchronize(handle_);
        }
        return status_;
    }

    /// Fused consumer: per-satellite envelope (min / max |position|, count of NONE results,
    /// first error and where) of `sat.propagate(times[k], sv)` over the whole grid, without
    /// materialising the states. `out` [size()], host or device memory; `jd_frac == nullptr`
    /// selects the propagate_from_epoch form with `jd` = minutes.
    int propagate_summary(
        const double *jd, const double *jd_frac, std::size_t m, perturb_gpu_summary *out,
        void *stream
    ) {
        status_ = perturb_gpu_propagate_summary(handle_, jd, jd_frac, m, out, stream);
        if (status_ == PERTURB_GPU_OK && stream == nullptr) {
            status_ = perturb_gpu_synchronize(handle_);
        }
        return status_;
    }

    perturb_gpu_batch *handle() { return handle_; }

private:
    SatelliteBatch() : handle_(nullptr), n_(0), status_(PERTURB_GPU_OK) {}

    static perturb_gpu_tle pack(const TwoLineElement &t) {
        perturb_gpu_tle p;
        p.epoch_year = static_cast<double>(t.epoch_year);
        p.epoch_day_of_year = t.epoch_day_of_year;
        p.n_dot = t.n_dot;
        p.n_ddot = t.n_ddot;
        p.b_star = t.b_star;
        p.inclination = t.inclination;
        p.raan = t.raan;
        p.eccentricity = t.eccentricity;
        p.arg_of_perigee = t.arg_of_perigee;
        p.mean_anomaly = t.mean_anomaly;
        p.mean_motion = t.mean_motion;
        return p;
    }

    static int grav_code(GravModel g) {
        switch (g) {
            case GravModel::WGS72_OLD: return PERTURB_GPU_WGS72_OLD;
            case GravModel::WGS84: return PERTURB_GPU_WGS84;
            default: return PERTURB_GPU_WGS72;
        }
    }

    void init(const std::vector<perturb_gpu_tle> &packed, GravModel grav_model, int device) {
        init_error_.assign(n_, static_cast<std::int32_t>(Sgp4Error::UNKNOWN));
        // 'i': the opsmode the reference's public constructors hard-wire (perturb.cpp:182,200)
        status_ = perturb_gpu_init(
            packed.data(), n_, grav_code(grav_model), 'i', device, &handle_, init_error_.data()
        );
        if (status_ != PERTURB_GPU_OK) {
            return;
        }
        fetch_queries();
    }

    void fetch_queries() {
        status_ = perturb_gpu_last_error(handle_, init_error_.data());
        epoch_jd_.assign(n_, 0.0);
        epoch_frac_.assign(n_, 0.0);
        if (status_ == PERTURB_GPU_OK) {
            status_ = perturb_gpu_epoch(handle_, epoch_jd_.data(), epoch_frac_.data());
        }
    }

    void scatter(
        const std::vector<double> &posvel, const std::vector<std::int32_t> &codes,
        std::size_t m, StateVector *svs, Sgp4Error *errs
    ) const {
        const std::size_t plane = n_ * m;
        for (std::size_t e = 0; e < plane; ++e) {
            const std::int32_t code = codes[e];
            if (errs != nullptr) {
                errs[e] = static_cast<Sgp4Error>(code);
            }
            if (code == 0 || code == 6) {  // the reference writes r, v for NONE and DECAYED only
                for (std::size_t c = 0; c < 3; ++c) {
                    svs[e].position[c] = posvel[c * plane + e];
                    svs[e].velocity[c] = posvel[(c + 3) * plane + e];
                }
            }
        }
    }

    perturb_gpu_batch *handle_;
    std::size_t n_;
    int status_;
    std::vector<std::int32_t> init_error_;
    std::vector<double> epoch_jd_, epoch_frac_;
};

}  // namespace perturb

#endif  // PERTURB_SATELLITE_BATCH_HPP
