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

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../perturb_elsetrec_layout.h"
#include "../perturb_gpu.h"
#include "batch_types.hpp"

namespace perturb {

/// Page-locks memory the caller already owns for the lifetime of this object (RAII over
/// perturb_gpu_host_register / _unregister), e.g. the storage of the `std::vector<StateVector>`
/// a propagate call fills: the copy engine then writes into it directly instead of going through
/// bounce buffers. Construct it after the vector has its final size and destroy it before the
/// vector is resized or freed.
class PinnedRegion {
public:
    PinnedRegion(void *p, std::size_t bytes)
        : p_(perturb_gpu_host_register(p, bytes) == PERTURB_GPU_OK ? p : nullptr) {}
    ~PinnedRegion() { perturb_gpu_host_unregister(p_); }
    PinnedRegion(const PinnedRegion &) = delete;
    PinnedRegion &operator=(const PinnedRegion &) = delete;
    bool ok() const { return p_ != nullptr; }

private:
    void *p_;
};

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
    /// `svs` and `errs` are satellite-major [size() * m], any host memory (`errs` may be null).
    /// The device writes the 64-byte `StateVector` records themselves; they reach `svs` through
    /// pinned bounce buffers and a few host threads (perturb_gpu_propagate_into). `sv.epoch` is
    /// always set and position / velocity are left untouched for error codes 1-4, exactly like
    /// the reference (sgp4.cpp:1865,1878,1943,1995).
    int propagate(const JulianDate *times, std::size_t m, StateVector *svs, Sgp4Error *errs) {
        std::vector<double> jd(m), frac(m);
        for (std::size_t k = 0; k < m; ++k) {
            jd[k] = times[k].jd;
            frac[k] = times[k].jd_frac;
        }
        return propagate_records(jd.data(), frac.data(), m, svs, errs);
    }

    /// Batched `Satellite::propagate_from_epoch(mins, sv)` (perturb.cpp:228-234);
    /// `sv.epoch = epoch(i) + mins / 1440` as the reference sets it.
    int propagate_from_epoch(
        const double *mins, std::size_t m, StateVector *svs, Sgp4Error *errs
    ) {
        return propagate_records(mins, nullptr, m, svs, errs);
    }

    /// The same without the save / restore: position and velocity of evaluations with error
    /// codes 1-4 are quiet NaN (the C ABI's convention). `svs` may be host or device memory.
    int propagate_states(
        const double *jd, const double *jd_frac, std::size_t m, StateVector *svs,
        Sgp4Error *errs, std::size_t row_stride, void *stream
    ) {
        status_ = perturb_gpu_propagate_states(
            handle_, jd, jd_frac, m, reinterpret_cast<perturb_gpu_state *>(svs),
            reinterpret_cast<std::int32_t *>(errs), row_stride, stream
        );
        if (status_ == PERTURB_GPU_OK && stream == nullptr) {
            status_ = perturb_gpu_synchronize(handle_);
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

    /// Batched `ClassicalOrbitalElements(sv, grav_model)` (perturb.cpp:119-131) over the planes
    /// a `propagate_soa` call wrote into DEVICE memory: eleven planes in the member order of
    /// `perturb::ClassicalOrbitalElements`, coe[e * coe_plane_stride + i * row_stride + k]
    /// (device memory; see perturb_gpu_rv2coe for sentinels and NaN conventions).
    static int classical_elements_soa(
        const double *posvel, std::size_t n, std::size_t m, std::size_t row_stride,
        std::size_t plane_stride, GravModel grav_model, double *coe,
        std::size_t coe_plane_stride, void *stream
    ) {
        return perturb_gpu_rv2coe(
            posvel, n, m, row_stride, plane_stride, grav_code(grav_model), coe, coe_plane_stride,
            stream
        );
    }

    /// Fills `records` -- an array of the reference's `sgp4::elsetrec`, or equally of
    /// `perturb::Satellite` (its only member is the public `sat_rec`, perturb.hpp:244), record
    /// i at byte `i * stride` -- with what `sgp4init` left for satellite i (every numeric
    /// member, error / isimp / irez, method, init, operationmode, t = 0). Identity members
    /// (satnum, classification, epochyr ...) are left as they are. Host memory.
    int export_records(void *records, std::size_t stride) {
        return status_ = perturb_gpu_export_records(handle_, records, stride);
    }

    /// What ONE `sats[i].propagate(t, sv)` leaves in `sats[i].sat_rec` besides the state: t,
    /// error, the mean elements am .. nm (sgp4.cpp:1801-1802, 1895-1901), for deep-space
    /// satellites aycof / xlcof / con41 / x1mth2 / x7thm1 and the resonance integrator state.
    int write_back_records(const JulianDate &t, void *records, std::size_t stride) {
        return status_ = perturb_gpu_write_back(handle_, t.jd, t.jd_frac, 1, records, stride);
    }
    /// The same for `propagate_from_epoch(mins, sv)`.
    int write_back_records_from_epoch(double mins, void *records, std::size_t stride) {
        return status_ = perturb_gpu_write_back(handle_, mins, 0.0, 0, records, stride);
    }

#ifdef PERTURB_PERTURB_HPP  // perturb's own headers are present: the typed forms
    /// `sats[i].sat_rec` <- the batch's initialised record, i < size().
    int export_records(Satellite *sats) { return export_records(sats, sizeof(Satellite)); }
    /// `sats[i].sat_rec` as `sats[i].propagate(t, sv)` would leave it, i < size().
    int write_back(Satellite *sats, const JulianDate &t) {
        return write_back_records(t, sats, sizeof(Satellite));
    }
    int write_back_from_epoch(Satellite *sats, double mins) {
        return write_back_records_from_epoch(mins, sats, sizeof(Satellite));
    }
#endif

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

#ifdef PERTURB_PERTURB_HPP
    // The generated layout table (include/perturb_elsetrec_layout.h) against the real struct:
    // a reference whose record moved does not compile instead of being written at wrong offsets.
    static_assert(sizeof(sgp4::elsetrec) == PERTURB_ELSETREC_SIZE, "elsetrec size moved");
    static_assert(sizeof(Satellite) == sizeof(sgp4::elsetrec), "Satellite is its sat_rec");
#  define PERTURB_BATCH_CHECK_MEMBER(name, offset, kind) \
    static_assert(offsetof(sgp4::elsetrec, name) == offset, "elsetrec member moved: " #name);
    PERTURB_ELSETREC_MEMBERS(PERTURB_BATCH_CHECK_MEMBER)
#  undef PERTURB_BATCH_CHECK_MEMBER
#endif

    // `StateVector` / `Sgp4Error` are handed to the C ABI as they are.
    static_assert(sizeof(StateVector) == sizeof(perturb_gpu_state), "StateVector is 64 bytes");
    static_assert(sizeof(JulianDate) == 2 * sizeof(double), "JulianDate is {jd, jd_frac}");
    static_assert(sizeof(Sgp4Error) == sizeof(std::int32_t), "Sgp4Error is an int enum");

    int propagate_records(
        const double *ta, const double *tb, std::size_t m, StateVector *svs, Sgp4Error *errs
    ) {
        if (n_ == 0 || m == 0) {
            return status_ = PERTURB_GPU_OK;
        }
        // perturb_gpu_propagate_into keeps the reference's semantics itself: epoch always set,
        // position / velocity stored for NONE and DECAYED only, any host memory as destination.
        return status_ = perturb_gpu_propagate_into(
                   handle_, ta, tb, m, reinterpret_cast<perturb_gpu_state *>(svs),
                   reinterpret_cast<std::int32_t *>(errs), m
               );
    }

    perturb_gpu_batch *handle_;
    std::size_t n_;
    int status_;
    std::vector<std::int32_t> init_error_;
    std::vector<double> epoch_jd_, epoch_frac_;
};

/// One process, several GPUs: the catalogue is split into contiguous satellite
/// ranges, one `SatelliteBatch` per device, with no exchange step -- every (satellite, time)
/// evaluation is independent and the time grid is replicated. Range boundaries balance the
/// estimated cost (deep-space evaluations cost 1.3-1.8x a near-Earth one), not the count.
/// A propagate call enqueues on every device first and synchronises afterwards, so with
/// pinned host buffers (`pinned_alloc`) all devices compute and copy into their slices of the
/// caller's arrays concurrently. Indices are the caller's, exactly as for `SatelliteBatch`.
/// `device_share` (optional, one number per device): the relative share of the cost each device
/// should get instead of an equal one -- e.g. the device -> host copy rate each GPU reaches while
/// all of them copy, when the results are gathered into host memory: that path is bound by the
/// host links, which the GPUs of a box do not always share evenly.
class ShardedSatelliteBatch {
public:
    ShardedSatelliteBatch(
        const TwoLineElement *tles, std::size_t n, const int *devices, std::size_t n_devices,
        GravModel grav_model = GravModel::WGS72, const double *device_share = nullptr
    )
        : n_(n), status_(PERTURB_GPU_OK) {
        if (n_devices == 0) {
            status_ = PERTURB_GPU_ERR_ARG;
            return;
        }
        std::vector<double> cost(n);
        for (std::size_t i = 0; i < n; ++i) {
            cost[i] = estimated_cost(tles[i]);
        }
        cuts_ = plan_ranges(cost, n_devices, device_share);
        for (std::size_t s = 0; s < n_devices; ++s) {
            shards_.emplace_back(
                tles + cuts_[s], cuts_[s + 1] - cuts_[s], grav_model, devices[s]
            );
            if (!shards_.back().ok()) {
                status_ = shards_.back().status();
            }
        }
    }

    /// Batched `Satellite::from_tle(line_1, line_2, grav)` over several devices: the lines are
    /// read once on the host (Vallado's numeric conventions, as `SatelliteBatch::from_tles`)
    /// to place the range boundaries, then every shard parses its own range on its device.
    static ShardedSatelliteBatch from_tles(
        const std::string *lines_1, const std::string *lines_2, std::size_t n, const int *devices,
        std::size_t n_devices, GravModel grav_model = GravModel::WGS72,
        const double *device_share = nullptr
    ) {
        ShardedSatelliteBatch out;
        out.n_ = n;
        if (n_devices == 0) {
            out.status_ = PERTURB_GPU_ERR_ARG;
            return out;
        }
        std::vector<double> cost(n);
        for (std::size_t i = 0; i < n; ++i) {
            perturb_gpu_tle numbers;
            const bool readable = lines_1[i].size() >= 69 && lines_2[i].size() >= 69 &&
                perturb_gpu_parse_tles(lines_1[i].c_str(), lines_2[i].c_str(), 70, 1, 1,
                                       &numbers, nullptr) == PERTURB_GPU_OK;
            TwoLineElement t = TwoLineElement();
            if (readable) {
                t.mean_motion = numbers.mean_motion;
                t.eccentricity = numbers.eccentricity;
            }
            cost[i] = estimated_cost(t);
        }
        out.cuts_ = plan_ranges(cost, n_devices, device_share);
        for (std::size_t s = 0; s < n_devices; ++s) {
            out.shards_.push_back(SatelliteBatch::from_tles(
                lines_1 + out.cuts_[s], lines_2 + out.cuts_[s], out.cuts_[s + 1] - out.cuts_[s],
                grav_model, devices[s]
            ));
            if (!out.shards_.back().ok()) {
                out.status_ = out.shards_.back().status();
            }
        }
        return out;
    }

    ShardedSatelliteBatch(ShardedSatelliteBatch &&) = default;
    ShardedSatelliteBatch &operator=(ShardedSatelliteBatch &&) = default;

    bool ok() const { return status_ == PERTURB_GPU_OK; }
    int status() const { return status_; }
    std::size_t size() const { return n_; }
    std::size_t shard_count() const { return shards_.size(); }
    /// Caller indices [begin, end) held by shard `s`.
    std::size_t shard_begin(std::size_t s) const { return cuts_[s]; }
    std::size_t shard_end(std::size_t s) const { return cuts_[s + 1]; }
    SatelliteBatch &shard(std::size_t s) { return shards_[s]; }

    Sgp4Error last_error(std::size_t i) const {
        const std::size_t s = shard_of(i);
        return s < shards_.size() ? shards_[s].last_error(i - cuts_[s]) : Sgp4Error::UNKNOWN;
    }
    JulianDate epoch(std::size_t i) const {
        const std::size_t s = shard_of(i);
        return s < shards_.size() ? shards_[s].epoch(i - cuts_[s]) : JulianDate();
    }

    /// The drop-in call, `Satellite::propagate(jd, sv)` (perturb.cpp:236-242) for every
    /// satellite and time into the caller's own arrays, satellite-major [size() * m]: every
    /// shard runs `SatelliteBatch::propagate` on its rows from its own host thread, so all
    /// devices compute, copy and scatter concurrently. Same semantics as the single-device
    /// class: `sv.epoch` always set, position / velocity untouched for error codes 1-4.
    int propagate(const JulianDate *times, std::size_t m, StateVector *svs, Sgp4Error *errs) {
        return for_each_shard([&](std::size_t s) {
            return shards_[s].propagate(
                times, m, svs + cuts_[s] * m, errs != nullptr ? errs + cuts_[s] * m : nullptr
            );
        });
    }

    /// `Satellite::propagate_from_epoch(mins, sv)` (perturb.cpp:228-234), likewise.
    int propagate_from_epoch(
        const double *mins, std::size_t m, StateVector *svs, Sgp4Error *errs
    ) {
        return for_each_shard([&](std::size_t s) {
            return shards_[s].propagate_from_epoch(
                mins, m, svs + cuts_[s] * m, errs != nullptr ? errs + cuts_[s] * m : nullptr
            );
        });
    }

    /// `SatelliteBatch::propagate_soa` over all shards into buffers indexed by the caller's
    /// satellite numbers: posvel[c * plane_stride + i * row_stride + k]. The buffers are
    /// either HOST memory (pinned recommended: every device copies its rows in concurrently)
    /// or memory of ONE device (`device_alloc`): the device-resident gather, where each
    /// shard's kernel stores its rows straight into that device over NVLink peer access.
    int propagate_soa(
        const double *jd, const double *jd_frac, std::size_t m, double *posvel,
        std::int32_t *err, std::size_t row_stride, std::size_t plane_stride
    ) {
        for (std::size_t s = 0; s < shards_.size() && status_ == PERTURB_GPU_OK; ++s) {
            if (cuts_[s + 1] == cuts_[s]) {
                continue;
            }
            const std::size_t off = cuts_[s] * row_stride;
            status_ = jd_frac != nullptr
                ? perturb_gpu_propagate(
                      shards_[s].handle(), jd, jd_frac, m, posvel + off, err + off, row_stride,
                      plane_stride, nullptr)
                : perturb_gpu_propagate_from_epoch(
                      shards_[s].handle(), jd, m, posvel + off, err + off, row_stride,
                      plane_stride, nullptr);
        }
        return synchronize();
    }

    /// `SatelliteBatch::propagate_states` over all shards (HOST records, NaN for codes 1-4);
    /// `jd_frac == nullptr` selects the propagate_from_epoch form with `jd` = minutes.
    int propagate_states(
        const double *jd, const double *jd_frac, std::size_t m, StateVector *svs,
        Sgp4Error *errs, std::size_t row_stride
    ) {
        for (std::size_t s = 0; s < shards_.size() && status_ == PERTURB_GPU_OK; ++s) {
            if (cuts_[s + 1] == cuts_[s]) {
                continue;
            }
            const std::size_t off = cuts_[s] * row_stride;
            status_ = perturb_gpu_propagate_states(
                shards_[s].handle(), jd, jd_frac, m,
                reinterpret_cast<perturb_gpu_state *>(svs + off),
                reinterpret_cast<std::int32_t *>(errs + off), row_stride, nullptr
            );
        }
        return synchronize();
    }

    int synchronize() {
        for (std::size_t s = 0; s < shards_.size(); ++s) {
            const int st = perturb_gpu_synchronize(shards_[s].handle());
            if (st != PERTURB_GPU_OK && status_ == PERTURB_GPU_OK) {
                status_ = st;
            }
        }
        return status_;
    }

    /// Pinned host memory every device can copy into asynchronously.
    static void *pinned_alloc(std::size_t bytes) {
        void *p = nullptr;
        return perturb_gpu_host_alloc(bytes, &p) == PERTURB_GPU_OK ? p : nullptr;
    }
    static void pinned_free(void *p) { perturb_gpu_host_free(p); }
    /// Device memory (the target of a device-resident gather) without linking the CUDA runtime.
    static void *device_alloc(int device, std::size_t bytes) {
        void *p = nullptr;
        return perturb_gpu_device_alloc(device, bytes, &p) == PERTURB_GPU_OK ? p : nullptr;
    }
    static void device_free(int device, void *p) { perturb_gpu_device_free(device, p); }
    /// Blocking copy between any two host / device buffers.
    static int copy(void *dst, const void *src, std::size_t bytes) {
        return perturb_gpu_copy(dst, src, bytes);
    }
    static int device_count() {
        int n = 0;
        return perturb_gpu_device_count(&n) == PERTURB_GPU_OK ? n : 0;
    }

    /// Relative cost of one evaluation of this TLE: the flops-v0 weight of the
    /// orbit class its numbers imply -- sgp4init's thresholds (sgp4.cpp:1590, 773-776, 1516)
    /// applied to the Kozai mean motion. Used to balance ranges only, never for results.
    static double estimated_cost(const TwoLineElement &t) {
        const double two_pi = 6.283185307179586;
        const double no = t.mean_motion * two_pi / 1440.0;  // rad/min
        if (!(no > 0.0)) {
            return 1170.0;
        }
        if (two_pi / no >= 225.0) {
            if (no > 0.0034906585 && no < 0.0052359877) {
                return 1586.0;  // 24 h resonant
            }
            if (no >= 8.26e-3 && no <= 9.24e-3 && t.eccentricity >= 0.5) {
                return 2097.0;  // 12 h resonant
            }
            return 1502.0;
        }
        // perigee below 220 km: simplified drag (a = (xke / no)^(2/3) in Earth radii)
        const double a = std::cbrt(0.0743669161 / no);
        return a * a * (1.0 - t.eccentricity) < 220.0 / 6378.135 + 1.0 ? 1090.0 : 1170.0;
    }

    /// Cut points (n_shards + 1 of them, first 0, last n) of the contiguous split whose summed
    /// costs are as equal -- or, with `share`, as proportional to the shards' shares -- as a
    /// prefix split allows.
    static std::vector<std::size_t> plan_ranges(
        const std::vector<double> &cost, std::size_t n_shards, const double *share = nullptr
    ) {
        const std::size_t n = cost.size();
        std::vector<double> cum(n + 1, 0.0);
        for (std::size_t i = 0; i < n; ++i) {
            cum[i + 1] = cum[i] + cost[i];
        }
        double share_total = 0.0, share_before = 0.0;
        for (std::size_t s = 0; share != nullptr && s < n_shards; ++s) {
            share_total += share[s] > 0.0 ? share[s] : 0.0;
        }
        std::vector<std::size_t> cuts(1, 0);
        for (std::size_t s = 1; s < n_shards; ++s) {
            double fraction = static_cast<double>(s) / static_cast<double>(n_shards);
            if (share_total > 0.0) {
                share_before += share[s - 1] > 0.0 ? share[s - 1] : 0.0;
                fraction = share_before / share_total;
            }
            const double target = cum[n] * fraction;
            std::size_t k = cuts.back();
            while (k < n && cum[k + 1] - target < target - cum[k]) {
                ++k;  // stop at the prefix closest to the target
            }
            cuts.push_back(k);
        }
        cuts.push_back(n);
        return cuts;
    }

private:
    ShardedSatelliteBatch() : n_(0), status_(PERTURB_GPU_OK) {}

    // fn(shard) on one host thread per non-empty shard; first failure wins.
    template <class Fn>
    int for_each_shard(Fn fn) {
        std::vector<int> st(shards_.size(), PERTURB_GPU_OK);
        std::vector<std::thread> threads;
        for (std::size_t s = 0; s < shards_.size(); ++s) {
            if (cuts_[s + 1] == cuts_[s]) {
                continue;
            }
            threads.emplace_back([&st, &fn, s]() { st[s] = fn(s); });
        }
        for (std::size_t t = 0; t < threads.size(); ++t) {
            threads[t].join();
        }
        for (std::size_t s = 0; s < st.size(); ++s) {
            if (st[s] != PERTURB_GPU_OK && status_ == PERTURB_GPU_OK) {
                status_ = st[s];
            }
        }
        return status_;
    }

    std::size_t shard_of(std::size_t i) const {
        for (std::size_t s = 0; s + 1 < cuts_.size(); ++s) {
            if (i >= cuts_[s] && i < cuts_[s + 1]) {
                return s;
            }
        }
        return shards_.size();
    }

    std::size_t n_;
    int status_;
    std::vector<std::size_t> cuts_;
    std::vector<SatelliteBatch> shards_;
};

}  // namespace perturb

#endif  // PERTURB_SATELLITE_BATCH_HPP
