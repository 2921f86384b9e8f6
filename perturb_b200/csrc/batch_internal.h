// Internal launch interface between the C-ABI (batch.cu) and the kernels.
#ifndef PERTURB_B200_BATCH_INTERNAL_H
#define PERTURB_B200_BATCH_INTERNAL_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "sgp4_fields.h"

namespace pb {

struct GravModelConsts;
struct TleNumbers;

// Records `msg` as the calling thread's perturb_gpu_strerror() text and returns `status`.
int report_failure(int status, const char *msg);

struct GravAll {  // host copy of getgravconst(), src/sgp4.cpp:2101-2157
    double tumin, mus, radiusearthkm, xke, j2, j3, j4, j3oj2;
};

// Kernel 1. tles: device AoS [n]; outputs: rec [RF_COUNT][n], pf_stage [PF_COUNT][n],
// cls [n], init_error [n] (raw sgp4 codes, 7 for an invalid TLE).
void launch_sgp4init(const void *d_tles, int n, const GravAll &g, bool opsmode_a, double *d_rec,
                     double *d_pf_stage, uint8_t *d_cls, int32_t *d_init_error,
                     cudaStream_t stream);

// TLE text -> TleNumbers on the device (one thread per TLE). Lines are fixed-stride records.
void launch_tle_parse(const char *d_lines1, const char *d_lines2, size_t stride, int n, int flavor,
                      void *d_tles, int32_t *d_status, cudaStream_t stream);

// Alternate entry: records already initialised (by the caller's own sgp4init); only derives
// the propagation fields and the class. init_error is taken from rec[RF_error].
void launch_derive_from_records(int n, const double *d_rec, double *d_pf_stage, uint8_t *d_cls,
                                int32_t *d_init_error, cudaStream_t stream);

// Gather caller-order staging into the class-sorted propagation layout.
// perm[slot] = caller index or -1 for padding slots.
void launch_pack(const double *d_pf_stage, int n, const int32_t *d_perm, int n_pad,
                 double *d_pf_sorted, cudaStream_t stream);

struct PropLaunch {
    const double *pf;  // [PF_COUNT][n_pad]
    size_t n_pad;
    const int32_t *perm;    // [n_pad]
    int tile_begin[PB_CLS_COUNT];  // first 32-satellite tile of each class
    int tile_count[PB_CLS_COUNT];
    const double *ta;  // jd[m] or mins[m]   (device)
    const double *tb;  // jd_frac[m] or null
    int m;
    int use_jd;
    double *out;  // out[c * plane_stride + sat * row_stride + k], c = x,y,z,vx,vy,vz
    size_t plane_stride, row_stride;
    int32_t *err;  // err[sat * err_row_stride + k]
    size_t err_row_stride;
    double *elements;  // optional 7 planes am, em, im, Om, om, mm, nm (same strides as out)
    // array-of-StateVector mode (out unused): 8 doubles per record at (sat * row_stride + k)
    double *states;
    // fused visibility screen (out/err unused): stations [S][8] doubles = Earth-fixed position
    // x, y, z [km], unit zenith ux, uy, uz, sin(minimum elevation), unused; gmst [m][2] = sin,
    // cos of the sidereal angle; scratch vis_words [n][S][chunks] (one bit per time), vis_best
    // [n][S][chunks] {largest sin(elevation), its time index}; results pass_out [n][S]
    // (perturb_gpu_pass_summary), passes [n][S][max_passes][2] or null
    const double *stations;
    int n_stations;
    const double *gmst;
    uint32_t *vis_words;
    double2 *vis_best;
    void *pass_out;
    int32_t *passes;
    int max_passes;
    // fused proximity screen: primaries [P][3][m] TEME km; scratch prox_partials [n][P][chunks]
    // {smallest squared distance, its time index, count below the threshold}; results
    // approach_out [n][P] (perturb_gpu_approach)
    const double *primaries;
    int n_primaries;
    double threshold_km;
    void *prox_partials;
    void *approach_out;
    // fused summary mode (out/err unused): partials [n][chunks], then summary [n]
    void *summary_partials;  // perturb_gpu_summary [n * summary_chunks], scratch
    void *summary_out;       // perturb_gpu_summary [n]
    size_t n;                // satellites (caller count), needed by the reduce pass
    double radiusearthkm, xke, j2, j3oj2;
    int opsmode_a;
    int cta_sats;  // satellites per CTA of kernel 2 (32 / 16 / 8; 0 = chosen from the grid size)
    // resonance node table scratch (null / 0: resonant evaluations integrate in-line)
    double2 *res_nodes;  // [resonant slots][res_cap] records {xli, xni, xndt, xnddt} (2 double2 each)
    int *res_k0, *res_cnt;
    double *res_range;   // [4]
    int res_cap;
    // optional fork/join resources so class launches overlap (null = serial on one stream)
    cudaStream_t aux[PB_CLS_COUNT];
    cudaEvent_t fork;
    bool fork_recorded;  // set by launch_sgp4_propagate
    cudaEvent_t join[PB_CLS_COUNT];
};

// Kernel 2, one launch per non-empty class. Returns the number of launches issued, or -1 when
// the grid (satellite tiles x time tiles) exceeds what one launch can cover.
int launch_sgp4_propagate(const PropLaunch &p, cudaStream_t stream);

// Planes of the write-back pass, wb[WB_COUNT][n] in caller order (error, stage and resonant
// are small integers stored as doubles).
enum {
    WB_T = 0, WB_ERROR, WB_STAGE, WB_AM, WB_EM, WB_IM, WB_OMEGA, WB_ARGP, WB_MM, WB_NM,
    WB_AYCOF, WB_XLCOF, WB_CON41, WB_X1MTH2, WB_X7THM1, WB_RESONANT, WB_ATIME, WB_XLI, WB_XNI,
    WB_COUNT
};

// One evaluation per satellite at the time (ta, tb) [JulianDate pair, or minutes with use_jd
// = 0 in `l`], side outputs only; d_cls [n] orbit class per satellite, caller order.
void launch_write_back(const PropLaunch &l, const uint8_t *d_cls, double ta, double tb,
                       double *d_wb, cudaStream_t stream);

// (sin, cos) of gstime_SGP4(jd[k] + jd_frac[k]) -> d_sincos[2 k], [2 k + 1].
void launch_gmst(const double *d_jd, const double *d_jd_frac, int m, double *d_sincos,
                 cudaStream_t stream);

// Number of per-satellite partial records the summary mode needs for a grid of m times.
size_t summary_chunks(size_t m);

// sizeof the proximity screen's per-chunk partial record.
size_t prox_partial_bytes();

}  // namespace pb

#endif
