// Kernel 2 (`sgp4`): the per-(satellite, time) evaluation -- the hot path that replaces a loop
// over Satellite::propagate(JulianDate, StateVector&) (src/perturb.cpp:236-242).
//
// Work decomposition:
//   * a CTA owns one tile of 32 class-sorted satellites x one tile of the time grid;
//   * the tile's per-satellite constants are read coalesced from the field-major layout
//     (32 consecutive satellites = one 256 B row per field) and staged in shared memory,
//     and so is the CTA's slice of the time grid;
//   * inside a warp, lanes are consecutive TIMES of ONE satellite: every class/isimp/irez
//     branch is warp-uniform, constants are shared-memory broadcasts, and a warp's results
//     are 512 contiguous bytes per output plane, so stores are full-line, 16 B per lane,
//     straight into the caller-order row (the sort permutation costs nothing on output);
//   * each lane evaluates TPL = 2 consecutive times per satellite to form those 16 B stores.
// FP64 pipe bound; no tensor cores (not a contraction).
#include "batch_internal.h"
#include "sgp4_eval.cuh"

namespace pb {

namespace {

constexpr int kWarps = 4;                      // warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kTpl = 2;                        // times per lane
constexpr int kTimeTile = kThreads * kTpl;     // times per CTA

template <int CLS>
struct TileFields {
    // Fields of the propagation layout a class reads: [0, kEndA) and [kBeginB, kEndB).
    static constexpr int kEndA = (CLS == PB_CLS_NEAR_FULL)   ? PF_FULL_END
                                 : (CLS == PB_CLS_NEAR_SIMP) ? PF_NEAR_END
                                                             : PF_COMMON_END;
    static constexpr int kBeginB = PF_FULL_END;
    static constexpr int kEndB = (CLS == PB_CLS_DEEP_NR)   ? PF_DEEP_NR_END
                                 : (CLS == PB_CLS_DEEP_R1) ? PF_DEEP_R1_END
                                 : (CLS == PB_CLS_DEEP_R2) ? PF_COUNT
                                                           : PF_FULL_END;
    static constexpr int kRows = kEndA + (kEndB - kBeginB);
    __host__ __device__ static constexpr int row(int f) {
        return f < kEndA ? f : f - kBeginB + kEndA;
    }
};

template <int CLS>
struct SmemConsts {  // constants of satellite `s` of the staged tile: broadcast LDS.64
    const double *base;
    __device__ __forceinline__ double operator()(int f) const {
        return base[TileFields<CLS>::row(f) * PB_TILE_SATS];
    }
};

struct PropParams {
    const double *pf;
    size_t n_pad;
    const int32_t *perm;
    int tile_begin;
    int n_time_tiles;
    const double *ta;
    const double *tb;
    int m;
    int use_jd;
    double *out;
    size_t plane_stride, row_stride;
    int32_t *err;
    size_t err_row_stride;
    int vec_ok;
    GravConsts g;
    int opsmode_a;
};

template <int CLS>
__global__ void __launch_bounds__(kThreads)
sgp4_prop_kernel(const PropParams p) {
    using TF = TileFields<CLS>;
    __shared__ double s_pf[TF::kRows * PB_TILE_SATS];
    __shared__ double s_ta[kTimeTile];
    __shared__ double s_tb[kTimeTile];
    __shared__ int s_perm[PB_TILE_SATS];

    const int ttile = blockIdx.x % p.n_time_tiles;
    const int stile = p.tile_begin + blockIdx.x / p.n_time_tiles;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t sat0 = static_cast<size_t>(stile) * PB_TILE_SATS;

    // ---- stage constants (coalesced 256 B rows) and this CTA's slice of the time grid ----
    for (int r = warp; r < TF::kRows; r += kWarps) {
        const int f = r < TF::kEndA ? r : r - TF::kEndA + TF::kBeginB;
        s_pf[r * PB_TILE_SATS + lane] = p.pf[static_cast<size_t>(f) * p.n_pad + sat0 + lane];
    }
    if (threadIdx.x < PB_TILE_SATS) s_perm[threadIdx.x] = p.perm[sat0 + threadIdx.x];
    const int k_base = ttile * kTimeTile;
    for (int i = threadIdx.x; i < kTimeTile; i += kThreads) {
        // Lanes past the end of the grid re-evaluate the last time (never stored): a made-up
        // time such as 0 would be ~1e9 minutes from epoch and spin the resonance integrator.
        const int k = min(k_base + i, p.m - 1);
        s_ta[i] = p.ta[k];
        s_tb[i] = p.use_jd ? p.tb[k] : 0.0;
    }
    __syncthreads();

    const int i0 = (warp * 32 + lane) * kTpl;  // this lane's first time within the tile
    const int k0 = k_base + i0;                // ... and within the grid
    if (k_base + warp * 32 * kTpl >= p.m) return;  // whole warp beyond the grid
    double ta[kTpl], tb[kTpl];
#pragma unroll
    for (int j = 0; j < kTpl; ++j) {
        ta[j] = s_ta[i0 + j];
        tb[j] = s_tb[i0 + j];
    }
    const ClassFlags no_flags = {false, false, 0};
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

#pragma unroll 1
    for (int s = 0; s < PB_TILE_SATS; ++s) {
        const int dst = s_perm[s];
        if (dst < 0) continue;  // padding slot (warp-uniform)
        const SmemConsts<CLS> k = {s_pf + s};

        double res[kTpl][6];
        int code[kTpl];
#pragma unroll
        for (int j = 0; j < kTpl; ++j) {
            // Satellite::propagate: tsince = ((jd - jd0) + (frac - frac0)) * 1440, grouping
            // as JulianDate::operator- has it (src/perturb.cpp:80-83, 237-238).
            double t;
            if (p.use_jd) {
                const double d = __dadd_rn(__dsub_rn(ta[j], k(PF_JD0)), __dsub_rn(tb[j], k(PF_JDF0)));
                t = __dmul_rn(d, 1440.0);
            } else {
                t = ta[j];
            }
            code[j] = sgp4_eval<CLS, false>(k, p.g, p.opsmode_a != 0, no_flags, t, res[j], nullptr);
            if (code[j] != 0 && code[j] != 6) {
#pragma unroll
                for (int c = 0; c < 6; ++c) res[j][c] = nan;
            }
        }

        // ---- stores: caller-order row `dst`, 16 B per lane per plane ----
        if (k0 < p.m) {
            double *row = p.out + static_cast<size_t>(dst) * p.row_stride + k0;
            int32_t *erow = p.err + static_cast<size_t>(dst) * p.err_row_stride + k0;
            if (p.vec_ok && k0 + 1 < p.m) {
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    *reinterpret_cast<double2 *>(row + c * p.plane_stride) =
                        make_double2(res[0][c], res[1][c]);
                }
                *reinterpret_cast<int2 *>(erow) = make_int2(code[0], code[1]);
            } else {
#pragma unroll
                for (int j = 0; j < kTpl; ++j) {
                    if (k0 + j < p.m) {
#pragma unroll
                        for (int c = 0; c < 6; ++c) row[c * p.plane_stride + j] = res[j][c];
                        erow[j] = code[j];
                    }
                }
            }
        }
    }
}

template <int CLS>
void launch_class(const PropLaunch &l, PropParams p, cudaStream_t stream, int &launches) {
    const int tiles = l.tile_count[CLS];
    if (tiles <= 0) return;
    p.tile_begin = l.tile_begin[CLS];
    const long long blocks = static_cast<long long>(tiles) * p.n_time_tiles;
    // grid.x limit is 2^31-1; at 1M satellites x 10k times this is 31250 * 40 blocks.
    sgp4_prop_kernel<CLS><<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(p);
    ++launches;
}

}  // namespace

int launch_sgp4_propagate(const PropLaunch &l, cudaStream_t stream) {
    if (l.m <= 0) return 0;
    PropParams p;
    p.pf = l.pf;
    p.n_pad = l.n_pad;
    p.perm = l.perm;
    p.tile_begin = 0;
    p.n_time_tiles = (l.m + kTimeTile - 1) / kTimeTile;
    p.ta = l.ta;
    p.tb = l.tb;
    p.m = l.m;
    p.use_jd = l.use_jd;
    p.out = l.out;
    p.plane_stride = l.plane_stride;
    p.row_stride = l.row_stride;
    p.err = l.err;
    p.err_row_stride = l.err_row_stride;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(l.out) % 16) == 0) && (l.plane_stride % 2 == 0) &&
               (l.row_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(l.err) % 8) == 0) &&
               (l.err_row_stride % 2 == 0);
    p.g.radiusearthkm = l.radiusearthkm;
    p.g.xke = l.xke;
    p.g.j2 = l.j2;
    p.g.j3oj2 = l.j3oj2;
    p.g.vkmpersec = l.radiusearthkm * l.xke / 60.0;
    p.opsmode_a = l.opsmode_a;
    int launches = 0;
    launch_class<PB_CLS_NEAR_FULL>(l, p, stream, launches);
    launch_class<PB_CLS_NEAR_SIMP>(l, p, stream, launches);
    launch_class<PB_CLS_DEEP_NR>(l, p, stream, launches);
    launch_class<PB_CLS_DEEP_R1>(l, p, stream, launches);
    launch_class<PB_CLS_DEEP_R2>(l, p, stream, launches);
    return launches;
}

}  // namespace pb
