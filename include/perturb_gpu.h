/*
 * perturb_gpu.h -- C ABI of the B200 batched SGP4/SDP4 propagator.
 *
 * Drop-in boundary for ONE path of gunvirranu/perturb: a loop over
 *     Sgp4Error Satellite::propagate(JulianDate, StateVector&)          src/perturb.cpp:236-242
 *     Sgp4Error Satellite::propagate_from_epoch(double, StateVector&)   src/perturb.cpp:228-234
 * on satellites built by
 *     Satellite::Satellite(const TwoLineElement&, GravModel)            src/perturb.cpp:135-186
 *     Satellite::from_tle(char*, char*, GravModel)                      src/perturb.cpp:189-205
 * The reference has no FFI layer of its own; these entry points are what a binding for that
 * path would bind (INTEGRATION.md shows the C++ shim a maintainer adds to perturb).
 *
 * Conventions kept from the reference (include/perturb/perturb.hpp:37-60, README.md:19-20):
 *   - nothing throws; every call returns a status. Library status (this header) is separate
 *     from the per-satellite / per-evaluation `Sgp4Error` codes, which are DATA;
 *   - the caller owns every input and output buffer; the library owns only the handle;
 *   - a handle is not re-entrant (like a `Satellite`); distinct handles are independent;
 *   - position [km] / velocity [km/s] in TEME are written for codes NONE(0) and DECAYED(6)
 *     only; for codes 1-4 the reference leaves the caller's StateVector untouched
 *     (src/sgp4.cpp:1865,1878,1943,1995) -- the batch writes quiet NaNs there.
 *
 * Buffers may be host or device pointers (detected with cudaPointerGetAttributes). Device
 * outputs are written directly by the kernel; host outputs go through a device staging
 * buffer owned by the handle and an asynchronous D2H copy (pinned host memory recommended).
 * A device output buffer may live on ANOTHER GPU of the same process (device-resident gather
 * of a sharded catalogue): the kernel then stores its satellites' rows into the
 * peer's memory over NVLink while it computes -- peer access is enabled on first use, and the
 * call fails with PERTURB_GPU_ERR_ARG if the two devices are not peers. A time grid that
 * lives on another device is copied to the batch's device first.
 * There is no CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef PERTURB_GPU_H
#define PERTURB_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Library status codes (negative = failure). Never an Sgp4Error. */
enum {
    PERTURB_GPU_OK = 0,
    PERTURB_GPU_ERR_ARG = -1,    /* null / inconsistent arguments                      */
    PERTURB_GPU_ERR_CUDA = -2,   /* a CUDA runtime call failed; see perturb_gpu_strerror */
    PERTURB_GPU_ERR_NODEV = -3,  /* no usable CUDA device                              */
    PERTURB_GPU_ERR_ALLOC = -4   /* host or device allocation failed                   */
};

/* perturb::GravModel, include/perturb/perturb.hpp:69-73 (same order). */
enum { PERTURB_GPU_WGS72_OLD = 0, PERTURB_GPU_WGS72 = 1, PERTURB_GPU_WGS84 = 2 };

/* perturb::Sgp4Error, include/perturb/perturb.hpp:50-60 (same values). */
enum {
    PERTURB_SGP4_NONE = 0,
    PERTURB_SGP4_MEAN_ELEMENTS = 1,
    PERTURB_SGP4_MEAN_MOTION = 2,
    PERTURB_SGP4_PERT_ELEMENTS = 3,
    PERTURB_SGP4_SEMI_LATUS_RECTUM = 4,
    PERTURB_SGP4_EPOCH_ELEMENTS_SUB_ORBITAL = 5,
    PERTURB_SGP4_DECAYED = 6,
    PERTURB_SGP4_INVALID_TLE = 7,
    PERTURB_SGP4_UNKNOWN = 8
};

/* Numeric members of perturb::TwoLineElement (include/perturb/tle.hpp:65-91), TLE units:
 * degrees, rev/day, two-digit epoch year + fractional day of year. Members the propagator
 * never reads (catalog number, designator, checksums ...) are not carried.
 * An entry whose epoch_year or mean_motion is NaN stands for a TLE the parser rejected: its
 * record is zeroed, its init error is INVALID_TLE and it propagates to MEAN_MOTION(2), which
 * is what the reference does with such a `Satellite` (src/perturb.cpp:190-196). */
typedef struct perturb_gpu_tle {
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
} perturb_gpu_tle;

typedef struct perturb_gpu_batch perturb_gpu_batch; /* opaque */

/* ---- construction -------------------------------------------------------------------- */

/* Replaces n x `Satellite(const TwoLineElement&, GravModel)` (src/perturb.cpp:135-186),
 * i.e. unit conversion + epoch + sgp4::sgp4init (src/sgp4.cpp:1383-1680) in kernel 1, then
 * classifies (near/deep, isimp, irez) and sorts the batch for kernel 2.
 *   tles       [n] host or device
 *   grav_model PERTURB_GPU_WGS72 is the reference default
 *   opsmode    'i' (what the reference's public API hard-wires, perturb.cpp:182,200) or 'a'
 *              (what its verification test uses, tests/test_perturb.cpp:35)
 *   device     CUDA device ordinal
 *   init_error [n] host, may be NULL: `Satellite::last_error()` after construction, caller order
 */
int perturb_gpu_init(const perturb_gpu_tle *tles, size_t n, int grav_model, char opsmode,
                     int device, perturb_gpu_batch **out, int32_t *init_error);

/* Replaces n x `Satellite::from_tle(line_1, line_2, grav)` (src/perturb.cpp:189-217) when
 * flavor = 1, or n x `TwoLineElement::parse` + `Satellite(tle, grav)` when flavor = 0, with the
 * TLE text parsed ON THE DEVICE (one thread per TLE; same code and therefore same bits as
 * perturb_gpu_parse_tles). lines1 / lines2: n fixed-stride records (host or device),
 * >= 69 characters used. parse_status [n] host, may be NULL (see perturb_gpu_parse_tles). */
int perturb_gpu_init_from_text(const char *lines1, const char *lines2, size_t line_stride,
                               size_t n, int flavor, int grav_model, char opsmode, int device,
                               perturb_gpu_batch **out, int32_t *init_error,
                               int32_t *parse_status);

/* Alternate entry: satellites whose `sgp4::elsetrec` was initialised elsewhere (replaces
 * n x `Satellite(sgp4::elsetrec)`, src/perturb.cpp:133). `records` is field-major
 * [perturb_gpu_record_field_count()][n] doubles (host or device), field order given by
 * perturb_gpu_record_field_name(). Kernel 1 is skipped; this isolates kernel 2. */
int perturb_gpu_init_records(const double *records, size_t n, int grav_model, char opsmode,
                             int device, perturb_gpu_batch **out, int32_t *init_error);

void perturb_gpu_free(perturb_gpu_batch *batch);

/* ---- the hot path -------------------------------------------------------------------- */

/* Replaces, for every satellite i < n and every time k < m,
 *     err = sats[i].propagate(JulianDate(jd[k], jd_frac[k]), sv)        perturb.cpp:236-242
 * Outputs (host or device, caller-owned, satellites in CALLER order):
 *   posvel  six planes x, y, z [km], vx, vy, vz [km/s]:
 *             posvel[c * plane_stride + i * row_stride + k]       (strides in doubles)
 *   err     err[i * row_stride + k], Sgp4Error values
 * row_stride >= m lets a caller fill a wider [n][M] matrix one time-chunk at a time;
 * plane_stride >= n * row_stride. `stream` is a cudaStream_t (NULL = default stream). The
 * call returns after enqueueing when all buffers are device or pinned-host memory; use
 * perturb_gpu_synchronize() before reading host outputs. */
int perturb_gpu_propagate(perturb_gpu_batch *batch, const double *jd, const double *jd_frac,
                          size_t m, double *posvel, int32_t *err, size_t row_stride,
                          size_t plane_stride, void *stream);

/* Same for `propagate_from_epoch(mins[k], sv)` (perturb.cpp:228-234). */
int perturb_gpu_propagate_from_epoch(perturb_gpu_batch *batch, const double *mins, size_t m,
                                     double *posvel, int32_t *err, size_t row_stride,
                                     size_t plane_stride, void *stream);

/* The same two calls writing `perturb::StateVector` records (include/perturb/perturb.hpp:
 * 198-205: JulianDate epoch {jd, jd_frac}; Vec3 position; Vec3 velocity -- 8 doubles, 64 bytes,
 * layout asserted in include/perturb/satellite_batch.hpp) instead of structure-of-arrays planes:
 * the buffer a caller's `std::vector<StateVector>` already is.
 *   states  states[i * row_stride + k] (row_stride in records, >= m), host or device memory
 *           (device memory 16-byte aligned), satellites in caller order. Pinned host memory is
 *           written by the copy engine and the call returns after enqueueing; pageable host
 *           memory is filled through the handle's bounce buffers before the call returns
 *   err     err[i * row_stride + k]
 * `epoch` is always set, as the reference does: to (jd[k], jd_frac[k]) by `propagate`
 * (perturb.cpp:240), to epoch(i) + mins[k] / 1440 (operator+, perturb.cpp:85-89: added to the
 * fractional part, not normalised) by `propagate_from_epoch` (perturb.cpp:229); jd_frac == NULL
 * selects that form with `jd` holding minutes. Position and velocity are quiet NaN for error
 * codes 1-4, where the reference leaves the caller's record untouched. */
typedef struct perturb_gpu_state {
    double epoch_jd, epoch_jd_frac;
    double position[3];
    double velocity[3];
} perturb_gpu_state;

int perturb_gpu_propagate_states(perturb_gpu_batch *batch, const double *jd,
                                 const double *jd_frac, size_t m, perturb_gpu_state *states,
                                 int32_t *err, size_t row_stride, void *stream);

/* The drop-in form of the loop itself,
 *     for i, k:  errs[i * row_stride + k] = sats[i].propagate(times[k], svs[i * row_stride + k]);
 * into the caller's own HOST arrays (any host memory: a std::vector<StateVector> is fine),
 * with the reference's semantics for failed evaluations: `epoch` is always set, position and
 * velocity are written for codes NONE and DECAYED only and otherwise LEFT AS THEY WERE
 * (src/sgp4.cpp:1865,1878,1943,1995 return before the store). `err` may be NULL. Synchronous:
 * the results are in place when the call returns. Chunks of the time grid are computed, copied
 * into pinned bounce buffers owned by the handle and moved into the caller's arrays by a few
 * host threads (PERTURB_GPU_HOST_THREADS, default half the hardware threads, at most 16) while
 * the next chunk is computed and copied. If `states` is page-locked memory (perturb_gpu_host_alloc,
 * or the caller's own array after perturb_gpu_host_register) the copy engine writes it directly
 * instead; the records of failed evaluations are saved beforehand and put back. jd_frac == NULL
 * selects the `propagate_from_epoch` form with `jd` holding minutes. */
int perturb_gpu_propagate_into(perturb_gpu_batch *batch, const double *jd, const double *jd_frac,
                               size_t m, perturb_gpu_state *states, int32_t *err,
                               size_t row_stride);

/* Same evaluation, additionally handing back what `sgp4()` leaves in `Satellite::sat_rec` on
 * every call (src/sgp4.cpp:1895-1901): the singly averaged mean elements
 *     elements[e * plane_stride + i * row_stride + k],  e = am, em, im, Om, om, mm, nm
 * (Earth radii, -, rad, rad, rad, rad, rad/min). They are written when the reference writes
 * them (codes 0, 3, 4, 6) and are quiet NaN otherwise (codes 1 and 2 return before the store).
 * jd_frac == NULL selects the `propagate_from_epoch` form with `jd` holding minutes.
 * posvel, err and elements are all device memory (asynchronous, written by the kernel) or all
 * host memory (time-chunked through device buffers of the call; returns when the results are
 * in place). */
int perturb_gpu_propagate_elements(perturb_gpu_batch *batch, const double *jd,
                                   const double *jd_frac, size_t m, double *posvel,
                                   int32_t *err, double *elements, size_t row_stride,
                                   size_t plane_stride, void *stream);

/* Fused consumer: per-satellite summary of a propagation computed on the
 * fly -- nothing of size [n][m] is written, which removes the 52 B/evaluation output (and the
 * PCIe transfer behind it) for screens that only need an envelope. Equivalent to running
 * `sat.propagate(t_k, sv)` for k = 0..m-1 and keeping, per satellite:
 *   r_min, k_min   smallest |sv.position| [km] over evaluations that returned NONE, and the
 *                  lowest k attaining it (r_min = +inf, k_min = -1 if there is none)
 *   r_max, k_max   largest |sv.position| likewise (r_max = -inf, k_max = -1 if none)
 *   n_ok           number of evaluations that returned NONE
 *   first_error    Sgp4Error of the lowest k that did not return NONE (0 if all did), and
 *   k_first_error  that k (-1 if none) -- where the reference's verification loop would stop
 *                  (tests/test_perturb.cpp:368-373)
 * jd_frac == NULL selects the `propagate_from_epoch` form with `jd` holding minutes.
 * `out` [n] may be host or device memory, satellites in caller order. */
typedef struct perturb_gpu_summary {
    double r_min, r_max;
    int32_t k_min, k_max;
    int32_t n_ok;
    int32_t first_error;
    int32_t k_first_error;
    int32_t reserved;
} perturb_gpu_summary;

int perturb_gpu_propagate_summary(perturb_gpu_batch *batch, const double *jd,
                                  const double *jd_frac, size_t m, perturb_gpu_summary *out,
                                  void *stream);

/* Fused consumer: ground-station visibility screen. Equivalent to running
 * `sat.propagate(JulianDate(jd[k], jd_frac[k]), sv)` for every satellite and time, rotating
 * sv.position from TEME into the Earth-fixed frame by the sidereal angle the reference's
 * gstime_SGP4(jd + jd_frac) gives (src/sgp4.cpp:2506-2525; polar motion ignored), and testing
 * the elevation seen from each station against its mask -- without materialising the states.
 * Evaluations that do not return NONE count as not visible. Per (satellite, station):
 *   n_visible          grid times at or above the mask
 *   n_passes           maximal runs of consecutive visible times
 *   k_first_rise       first visible index (-1 if none), k_last_visible likewise the last
 *   max_sin_elevation  largest sin(elevation) over all NONE evaluations, visible or not, and
 *   k_max_elevation    the lowest k attaining it (-inf / -1 if no evaluation returned NONE)
 * and, if `passes` is not NULL, the first `max_passes` passes as (rise index, set index) pairs
 * (set = last visible index of the pass; unused slots -1):
 *   passes[((i * n_stations + s) * max_passes + j) * 2 + {0, 1}]
 * `out` [n][n_stations] and `passes` may be host or device memory (both the same); stations
 * [n_stations <= 64] is HOST memory. Station positions are geodetic on the ellipsoid of the
 * batch's gravity model (flattening 1/298.26 for WGS-72, 1/298.257223563 for WGS-84). */
typedef struct perturb_gpu_station {
    double latitude_rad, longitude_rad, altitude_km, min_elevation_rad;
} perturb_gpu_station;

typedef struct perturb_gpu_pass_summary {
    double max_sin_elevation;
    int32_t k_max_elevation;
    int32_t n_visible;
    int32_t n_passes;
    int32_t k_first_rise;
    int32_t k_last_visible;
    int32_t reserved;
} perturb_gpu_pass_summary;

int perturb_gpu_propagate_visibility(perturb_gpu_batch *batch, const double *jd,
                                     const double *jd_frac, size_t m,
                                     const perturb_gpu_station *stations, size_t n_stations,
                                     perturb_gpu_pass_summary *out, int32_t *passes,
                                     size_t max_passes, void *stream);

/* Fused consumer: closest approach to a few primary objects (conjunction pre-filter). For every
 * satellite i and primary p, over the evaluations that return NONE:
 *   d_min_km  smallest |sv.position - primary_p(k)| and k_min, the lowest k attaining it
 *             (+inf / -1 if there is none)
 *   n_below   number of grid times closer than threshold_km
 * primaries: TEME positions [km] of the primaries ON THE SAME GRID, three planes per primary,
 *   primaries[(p * 3 + c) * m + k], c = x, y, z   (host or device memory; e.g. the position
 *   planes a perturb_gpu_propagate call wrote for a batch of the primaries; NaN = no state);
 * out [n][n_primaries] host or device memory, n_primaries <= 64. jd_frac == NULL selects the
 * `propagate_from_epoch` form with `jd` holding minutes. */
typedef struct perturb_gpu_approach {
    double d_min_km;
    int32_t k_min;
    int32_t n_below;
} perturb_gpu_approach;

int perturb_gpu_propagate_proximity(perturb_gpu_batch *batch, const double *jd,
                                    const double *jd_frac, size_t m, const double *primaries,
                                    size_t n_primaries, double threshold_km,
                                    perturb_gpu_approach *out, void *stream);

/* Replaces `ClassicalOrbitalElements(StateVector, GravModel)` (src/perturb.cpp:119-131, i.e.
 * sgp4::rv2coe_SGP4, src/sgp4.cpp:2863-3064) for every state of a propagate result -- what the
 * reference's verification printout does after each propagate (tests/test_perturb.cpp:385-397).
 *   posvel  the six planes written by perturb_gpu_propagate (DEVICE memory, same strides)
 *   coe     eleven planes coe[e * coe_plane_stride + i * row_stride + k] (DEVICE memory),
 *           e = semilatus_rectum [km], semimajor_axis [km], eccentricity, inclination, raan,
 *           arg_of_perigee, true_anomaly, mean_anomaly, arg_of_latitude, true_longitude,
 *           longitude_of_periapsis [rad]: the member order of perturb::ClassicalOrbitalElements
 *           (include/perturb/perturb.hpp:211-222). Elements the reference leaves "undefined"
 *           carry its 999999.1 sentinel; states that are NaN (error codes 1-4) give NaN. */
int perturb_gpu_rv2coe(const double *posvel, size_t n, size_t m, size_t row_stride,
                       size_t plane_stride, int grav_model, double *coe,
                       size_t coe_plane_stride, void *stream);

int perturb_gpu_synchronize(perturb_gpu_batch *batch);

/* ---- multi-GPU plumbing (one process, several devices) ------------------- */

/* Number of usable CUDA devices. */
int perturb_gpu_device_count(int *count);

/* Page-locked host memory, visible to every device (cudaHostAllocPortable): with pinned
 * output buffers a propagate call returns after enqueueing, so shards on several devices run
 * and copy into their slices of one buffer concurrently. */
int perturb_gpu_host_alloc(size_t bytes, void **out);
void perturb_gpu_host_free(void *p);

/* Page-locks memory the caller ALREADY owns (cudaHostRegister) -- e.g. the storage of a
 * std::vector<StateVector> -- so that the copy engine writes results into it directly:
 * perturb_gpu_propagate_into then needs no bounce copy (about twice the rate on a box whose host
 * memory bandwidth bounds the bounce path). Unregister BEFORE the memory is freed or reallocated.
 * Registering costs about as much as ten propagate calls into the same memory: worthwhile for
 * arrays that are reused. */
int perturb_gpu_host_register(void *p, size_t bytes);
void perturb_gpu_host_unregister(void *p);

/* Device memory for callers that do not link the CUDA runtime themselves: buffers for
 * device-resident results (and the target of a device-resident gather: shards on other GPUs
 * store into it over NVLink peer access). perturb_gpu_copy moves `bytes` between any two of
 * host / device buffers (cudaMemcpyDefault) and returns when the copy is complete. */
int perturb_gpu_device_alloc(int device, size_t bytes, void **out);
void perturb_gpu_device_free(int device, void *p);
int perturb_gpu_copy(void *dst, const void *src, size_t bytes);

/* ---- queries ------------------------------------------------------------------------- */

size_t perturb_gpu_size(const perturb_gpu_batch *batch);
int perturb_gpu_device(const perturb_gpu_batch *batch);

/* `Satellite::last_error()` right after construction, caller order; out [n] host. */
int perturb_gpu_last_error(const perturb_gpu_batch *batch, int32_t *out);

/* `Satellite::epoch()` (perturb.cpp:224-226); out arrays [n] host. */
int perturb_gpu_epoch(const perturb_gpu_batch *batch, double *jd, double *jd_frac);

/* Orbit class per satellite (0 near full drag, 1 near simplified, 2 deep non-resonant,
 * 3 deep 24 h resonant, 4 deep 12 h resonant); out [n] host. */
int perturb_gpu_orbit_class(const perturb_gpu_batch *batch, uint8_t *out);

/* The device-resident mirror of `Satellite::sat_rec` (sgp4::elsetrec numerics). */
int perturb_gpu_record_field_count(void);
const char *perturb_gpu_record_field_name(int field);
/* out [n] host, caller order. */
int perturb_gpu_record_field(const perturb_gpu_batch *batch, int field, double *out);

/* ---- write-back into the reference's own records ---------------------- */

/* sizeof(perturb::sgp4::elsetrec) == sizeof(perturb::Satellite) the layout table was generated
 * for (include/perturb_elsetrec_layout.h, from include/perturb/sgp4.hpp:89-134): 1000. */
int perturb_gpu_elsetrec_size(void);

/* Fills an array of the reference's `sgp4::elsetrec` (equally: of `perturb::Satellite`, whose
 * only member is the public `sat_rec`, include/perturb/perturb.hpp:244) with the state
 * `sgp4init` left: every numeric member the batch carries, error / isimp / irez, method, init
 * and operationmode, t = 0. Members that only identify the object (satnum, classification,
 * intldesg, epochyr, epochdays, ephtype, elnum, revnum ...) are NOT touched: build the records
 * with the reference's constructor, or zero them, first. `records` is HOST memory,
 * record i at byte i * stride, stride >= perturb_gpu_elsetrec_size(); caller order. */
int perturb_gpu_export_records(const perturb_gpu_batch *batch, void *records, size_t stride);

/* What ONE call `sats[i].propagate(JulianDate(jd, jd_frac), sv)` (use_jd != 0) or
 * `sats[i].propagate_from_epoch(jd, sv)` (use_jd == 0, `jd` = minutes) leaves in `sat_rec`
 * besides the state it returns, for every satellite i: t and error (src/sgp4.cpp:1801-1802),
 * the singly averaged mean elements am, em, im, Om, om, mm, nm where the reference stores them
 * (:1895-1901; not for codes 1 and 2), for deep-space satellites aycof, xlcof (:1952-1957) and
 * con41, x1mth2, x7thm1 (:2017-2019), and for resonant ones the integrator state atime, xli,
 * xni (:1143-1147). The reference's results never depend on that state (restarted and resumed
 * integrator walks are bit-identical), so this is for callers that READ sat_rec. Same
 * `records` conventions as perturb_gpu_export_records; synchronous. */
int perturb_gpu_write_back(perturb_gpu_batch *batch, double jd, double jd_frac, int use_jd,
                           void *records, size_t stride);

/* Number of kernel launches the last propagate call on this handle issued. */
int perturb_gpu_last_launch_count(const perturb_gpu_batch *batch);

/* Message for the most recent failure on the calling thread. */
const char *perturb_gpu_strerror(void);

/* ---- host-side TLE text front end ------------------------------------------------------ */

/* Fixed-column reader for n TLEs. lines1/lines2 hold n records of `line_stride` bytes each
 * (>= 69 used; no terminator needed). flavor 0 follows TwoLineElement::parse's numeric
 * conventions (src/tle.cpp:123-138), flavor 1 follows twoline2rv's (src/sgp4.cpp:2231-2338);
 * they differ in the last bit of n_ddot / b_star only. status [n] may be NULL: 0 ok, else a
 * perturb::TLEParseError value (include/perturb/tle.hpp:45-51); rejected entries are set to
 * NaN so that perturb_gpu_init() reports INVALID_TLE for them. */
int perturb_gpu_parse_tles(const char *lines1, const char *lines2, size_t line_stride, size_t n,
                           int flavor, perturb_gpu_tle *out, int32_t *status);

/* The members of perturb::TwoLineElement (include/perturb/tle.hpp:65-91) that are not numbers
 * the propagator reads. */
typedef struct perturb_gpu_tle_ident {
    char catalog_number[6];
    char classification;
    unsigned int launch_year;
    unsigned int launch_number;
    char launch_piece[4];
    unsigned char ephemeris_type;
    unsigned int element_set_number;
    unsigned char line_1_checksum;
    unsigned long revolution_number;
    unsigned char line_2_checksum;
} perturb_gpu_tle_ident;

/* One `TwoLineElement::parse(line_1, line_2)` (src/tle.cpp:41-173): returns the
 * perturb::TLEParseError value (0 NONE .. 4 CHECKSUM_MISMATCH, same checking order), or a
 * negative library status for null arguments. The reference's sscanf directives are restated
 * token for token, so blank or misplaced fields are rejected (or accepted) exactly as there.
 * `numbers` / `ident` are written when every field could be read (results 0, 3 and 4);
 * `ident` may be NULL. Both lines must hold 69 characters. */
int perturb_gpu_parse_tle(const char *line_1, const char *line_2, perturb_gpu_tle *numbers,
                          perturb_gpu_tle_ident *ident);

#ifdef __cplusplus
}
#endif
#endif /* PERTURB_GPU_H */
