// Kernel 1 (`sgp4init`) and the layout helpers around it. Compiled with -fmad=false, see
// sgp4_init.cuh. One thread per satellite; field-major (SoA) stores are coalesced.
#include "batch_internal.h"
#include "sgp4_init.cuh"

namespace pb {

namespace {

struct LocalConsts {  // propagation fields held in registers / local memory
    const double *pf;
    __device__ __forceinline__ double operator()(int f) const { return pf[f]; }
};

// The trailing `sgp4(satrec, 0.0, r, v)` of sgp4init (src/sgp4.cpp:1673): sets the error the
// caller sees from last_error() and leaves the epoch mean elements in the record.
__device__ void epoch_probe(SatRec &rec, const double pf[PF_COUNT], bool opsmode_a) {
    GravConsts g;
    g.radiusearthkm = rec.f[RF_radiusearthkm];
    g.xke = rec.f[RF_xke];
    g.j2 = rec.f[RF_j2];
    g.j3oj2 = rec.f[RF_j3oj2];
    g.vkmpersec = g.radiusearthkm * g.xke / 60.0;
    g.inv_xke = 1.0 / g.xke;
    ClassFlags fl;
    fl.deep = rec.f[RF_method_d] != 0.0;
    fl.isimp = rec.f[RF_isimp] != 0.0;
    fl.irez = static_cast<int>(rec.f[RF_irez]);
    double out[6];
    Sgp4Side side;
    const int code = sgp4_eval<PB_CLS_RUNTIME, true, false>(LocalConsts {pf}, g, opsmode_a, fl, 0.0,
                                                     out, &side);
    rec.f[RF_error] = code;
    if (side.stage >= 1) {
        rec.f[RF_am] = side.am;
        rec.f[RF_em] = side.em;
        rec.f[RF_im] = side.im;
        rec.f[RF_Om] = side.Om;
        rec.f[RF_om] = side.om;
        rec.f[RF_mm] = side.mm;
        rec.f[RF_nm] = side.nm;
    }
    if (fl.deep && side.stage >= 2) {
        rec.f[RF_aycof] = side.aycof;
        rec.f[RF_xlcof] = side.xlcof;
    }
    if (fl.deep && side.stage >= 3) {
        rec.f[RF_con41] = side.con41;
        rec.f[RF_x1mth2] = side.x1mth2;
        rec.f[RF_x7thm1] = side.x7thm1;
    }
}

__global__ void __launch_bounds__(128)
sgp4init_kernel(const TleNumbers *__restrict__ tles, int n, GravModelConsts gc, int opsmode_a,
                double *__restrict__ rec_out, double *__restrict__ pf_out,
                uint8_t *__restrict__ cls_out, int32_t *__restrict__ err_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TleNumbers t = tles[i];

    SatRec rec;
#pragma unroll
    for (int f = 0; f < RF_COUNT; ++f) rec.f[f] = 0.0;
    double pf[PF_COUNT];

    // An all-NaN entry marks a TLE the host parser rejected: the reference then holds a
    // zeroed record with error = INVALID_TLE (src/perturb.cpp:190-196).
    const bool valid = !isnan(t.epoch_year) && !isnan(t.mean_motion);
    if (valid) {
        // unit conversion and epoch, src/perturb.cpp:136-178
        constexpr double deg2rad = kPi / 180.0;
        constexpr double xpdotp = 1440.0 / (2 * kPi);
        const double no_kozai = t.mean_motion / xpdotp;
        const double ndot = t.n_dot / (xpdotp * 1440.0);
        const double nddot = t.n_ddot / (xpdotp * 1440.0 * 1440);
        double jd, jdf;
        epoch_to_jd(t.epoch_year, t.epoch_day_of_year, jd, jdf);
        rec.f[RF_jdsatepoch] = jd;
        rec.f[RF_jdsatepochF] = jdf;
        const double epoch = (jd + jdf) - 2433281.5;
        sgp4init_core(rec, gc, epoch, t.b_star, ndot, nddot, t.eccentricity,
                      t.arg_of_perigee * deg2rad, t.inclination * deg2rad,
                      t.mean_anomaly * deg2rad, no_kozai, t.raan * deg2rad);
        derive_prop_fields(rec, pf);
        epoch_probe(rec, pf, opsmode_a != 0);
    } else {
        derive_prop_fields(rec, pf);
        pf[PF_AOBASE] = 0.0;  // pow(0/0) would be NaN; nm == 0 exits with code 2 first anyway
        rec.f[RF_error] = 7.0;
    }

    const size_t stride = static_cast<size_t>(n);
#pragma unroll
    for (int f = 0; f < RF_COUNT; ++f) rec_out[f * stride + i] = rec.f[f];
#pragma unroll
    for (int f = 0; f < PF_COUNT; ++f) pf_out[f * stride + i] = pf[f];
    cls_out[i] = static_cast<uint8_t>(class_of(rec));
    err_out[i] = static_cast<int32_t>(rec.f[RF_error]);
}

__global__ void __launch_bounds__(128)
derive_kernel(int n, const double *__restrict__ rec_in, double *__restrict__ pf_out,
              uint8_t *__restrict__ cls_out, int32_t *__restrict__ err_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t stride = static_cast<size_t>(n);
    SatRec rec;
#pragma unroll
    for (int f = 0; f < RF_COUNT; ++f) rec.f[f] = rec_in[f * stride + i];
    double pf[PF_COUNT];
    derive_prop_fields(rec, pf);
    if (!(rec.f[RF_no_unkozai] > 0.0)) pf[PF_AOBASE] = 0.0;
#pragma unroll
    for (int f = 0; f < PF_COUNT; ++f) pf_out[f * stride + i] = pf[f];
    cls_out[i] = static_cast<uint8_t>(class_of(rec));
    err_out[i] = static_cast<int32_t>(rec.f[RF_error]);
}

// One thread per (slot, field-group): gathers a satellite's staged fields into its sorted slot.
__global__ void __launch_bounds__(256)
pack_kernel(const double *__restrict__ stage, int n, const int32_t *__restrict__ perm, int n_pad,
            double *__restrict__ sorted) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (slot >= n_pad) return;
    const int src = perm[slot];
    // Padding slots replicate nothing: zeros evaluate to error 2 and are never stored.
    const double v = (src >= 0) ? stage[static_cast<size_t>(f) * n + src] : 0.0;
    sorted[static_cast<size_t>(f) * n_pad + slot] = v;
}

}  // namespace

// Pre-pass of the visibility screen: (sin, cos) of the Greenwich mean sidereal angle of every
// grid time, theta = gstime_SGP4(jd + jd_frac) (src/sgp4.cpp:2506-2525; the rotation about z
// that takes TEME to the Earth-fixed frame, polar motion ignored). In this translation unit
// (-fmad=false) the angle is the reference formula's double bit for bit.
__global__ void __launch_bounds__(256)
gmst_kernel(const double *__restrict__ jd, const double *__restrict__ jd_frac, int m,
            double *__restrict__ sincos_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const double theta = gstime_dev(jd[k] + jd_frac[k]);
    double s, c;
    sincos(theta, &s, &c);
    sincos_out[2 * k] = s;
    sincos_out[2 * k + 1] = c;
}

void launch_gmst(const double *d_jd, const double *d_jd_frac, int m, double *d_sincos,
                 cudaStream_t stream) {
    if (m <= 0) return;
    gmst_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_jd, d_jd_frac, m, d_sincos);
}

void launch_sgp4init(const void *d_tles, int n, const GravAll &g, bool opsmode_a, double *d_rec,
                     double *d_pf_stage, uint8_t *d_cls, int32_t *d_init_error,
                     cudaStream_t stream) {
    if (n <= 0) return;
    GravModelConsts gc;
    gc.tumin = g.tumin;
    gc.mus = g.mus;
    gc.radiusearthkm = g.radiusearthkm;
    gc.xke = g.xke;
    gc.j2 = g.j2;
    gc.j3 = g.j3;
    gc.j4 = g.j4;
    gc.j3oj2 = g.j3oj2;
    const int threads = 128;
    const int blocks = (n + threads - 1) / threads;
    sgp4init_kernel<<<blocks, threads, 0, stream>>>(static_cast<const TleNumbers *>(d_tles), n,
                                                    gc, opsmode_a ? 1 : 0, d_rec, d_pf_stage,
                                                    d_cls, d_init_error);
}

void launch_derive_from_records(int n, const double *d_rec, double *d_pf_stage, uint8_t *d_cls,
                                int32_t *d_init_error, cudaStream_t stream) {
    if (n <= 0) return;
    const int threads = 128;
    const int blocks = (n + threads - 1) / threads;
    derive_kernel<<<blocks, threads, 0, stream>>>(n, d_rec, d_pf_stage, d_cls, d_init_error);
}

void launch_pack(const double *d_pf_stage, int n, const int32_t *d_perm, int n_pad,
                 double *d_pf_sorted, cudaStream_t stream) {
    if (n_pad <= 0) return;
    const dim3 block(256);
    const dim3 grid((n_pad + 255) / 256, PF_COUNT);
    pack_kernel<<<grid, block, 0, stream>>>(d_pf_stage, n, d_perm, n_pad, d_pf_sorted);
}

}  // namespace pb
