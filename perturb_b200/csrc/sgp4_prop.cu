// Kernel 2 (`sgp4`): the per-(satellite, time) evaluation -- the hot path that replaces a loop
// over Satellite::propagate(JulianDate, StateVector&) (src/perturb.cpp:236-242).
//
// Work decomposition:
//   * a CTA owns 32, 16 or 8 satellites of one tile of 32 class-sorted satellites x one tile of
//     the time grid (fewer satellites per CTA on launches of few waves: shorter last wave);
//   * the tile's per-satellite constants are read coalesced from the field-major layout
//     (32 consecutive satellites = one 256 B row per field) and staged in shared memory,
//     and so is the CTA's slice of the time grid;
//   * inside a warp, lanes are consecutive TIMES of ONE satellite: every class/isimp/irez
//     branch is warp-uniform, constants are shared-memory broadcasts, and one evaluation of
//     a warp yields 256 contiguous bytes per output plane, stored as full sectors straight
//     into the caller-order row (the sort permutation costs nothing on output);
//   * each lane evaluates PB_TPL times per satellite: 1 (default: a warp = 32 consecutive times, a
//     CTA = 32 satellites x 128 times) or 2 (l and l + 32 of the warp's 64; measured equal on
//     long grids and 3.6 % slower on 1440 times, whose last half-filled 64-group it wastes).
// FP64 pipe bound; no tensor cores (not a contraction).
#include "../../include/perturb_gpu.h"
#include "batch_internal.h"
#include "sgp4_eval.cuh"

namespace pb {

namespace {

// Resident CTAs per SM the register allocation is capped for (measured on B200, sweeps in
// profiles/): the near-Earth bodies need 64-72 registers and do best at 7 CTAs (28 warps/SM;
// 6 and 8 are within 1.5 %), the deep-space bodies at 6 (80 registers: the non-resonant and 24 h
// bodies fit, the 12 h-resonant one spills 92 bytes and is level with 5 CTAs at 96 registers
// without spills, profiles/r2m_variants.jsonl).
#ifndef PB_MIN_BLOCKS_NEAR
#define PB_MIN_BLOCKS_NEAR 7
#endif
#ifndef PB_MIN_BLOCKS_DEEP
#define PB_MIN_BLOCKS_DEEP 6
#endif
#ifndef PB_MIN_BLOCKS_R2
#define PB_MIN_BLOCKS_R2 6
#endif
#ifndef PB_WARPS
#define PB_WARPS 4
#endif
constexpr int kWarps = PB_WARPS;               // warps per CTA
constexpr int kThreads = kWarps * 32;
#ifndef PB_TPL
#define PB_TPL 1
#endif
constexpr int kTpl = PB_TPL;                   // times per lane (1 or 2)
constexpr int kTimeTile = kThreads * kTpl;     // times per CTA
// PB_TPL = 2 comes in two forms. Rolled (default): lane l owns times l and l + 32 of the warp's
// 64, one copy of the body. PB_TPL_PAIR: lane l owns the ADJACENT times 2l and 2l + 1, the two
// evaluations are unrolled into one instruction stream (constants, addresses and loop control
// shared, two independent dependency chains) and the state planes are written with 16-byte
// stores, 512 contiguous bytes per warp and plane.
#ifndef PB_TPL_PAIR
#define PB_TPL_PAIR 0
#endif
#ifndef PB_SMEM_VECTOR_ADDR
#define PB_SMEM_VECTOR_ADDR 1
#endif
constexpr bool kPair = (PB_TPL_PAIR != 0) && (kTpl == 2);

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
        return base[TileFields<CLS>::row(f)];
    }
};

struct ProxPartial {  // proximity screen, one warp's 32 times of one (satellite, primary)
    double d2_min;  // smallest squared distance over NONE evaluations (+inf if none)
    int k_min;      // its time index (-1 if none)
    int n_below;    // evaluations closer than the threshold
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
    double *elements;
    double *states;  // perturb_gpu_state [n][row_stride] as doubles (AoS mode)
    perturb_gpu_summary *partials;  // summary mode: [sat * chunks + chunk]
    int chunks;
    // visibility screen
    const double *stations;  // [S][8]
    int n_stations;
    const double *gmst;      // [m][2]
    uint32_t *vis_words;     // [(sat * S + station) * chunks + chunk]
    double2 *vis_best;       // same index: {largest sin(elevation), its time index}
    // proximity screen
    const double *primaries;  // [P][3][m]
    int n_primaries;
    double thr2;              // squared threshold [km^2]
    ProxPartial *prox;        // [(sat * P + primary) * chunks + chunk]
    GravConsts g;
    int opsmode_a;
    int vec_ok;  // state planes and error rows allow 16-byte / 8-byte paired stores (kPair)
    // resonance node table (deep-space resonant classes only)
    const ResNode *res_nodes;  // [resonant slot][res_cap] node records
    const int *res_k0;         // [resonant slots] first tabulated step index
    const int *res_cnt;        // [resonant slots] number of tabulated nodes
    int res_cap;               // records per resonant slot
    size_t res_first_slot;     // sorted slot of the first resonant satellite
};

struct GlobalConsts {  // constants of one sorted slot straight from the field-major layout
    const double *base;
    size_t stride;
    __device__ __forceinline__ double operator()(int f) const { return base[f * stride]; }
};

// tsince of a grid time for one satellite: Satellite::propagate, src/perturb.cpp:237-238 with
// JulianDate::operator- (:80-83); propagate_from_epoch passes minutes through.
__device__ __forceinline__ double tsince_of(int use_jd, double ta, double tb, double jd0,
                                            double jdf0) {
    if (!use_jd) return ta;
    return __dmul_rn(__dadd_rn(__dsub_rn(ta, jd0), __dsub_rn(tb, jdf0)), 1440.0);
}

// Pre-pass 1: earliest and latest time of the grid -> range[0..3] = (ta, tb) of the earliest,
// (ta, tb) of the latest. One CTA; the grid is at most a few 10^5 entries.
__global__ void __launch_bounds__(256)
time_range_kernel(const double *__restrict__ ta, const double *__restrict__ tb, int m, int use_jd,
                  double *__restrict__ range) {
    __shared__ double s_lo[256], s_hi[256];
    __shared__ int s_ilo[256], s_ihi[256];
    const double ref_a = ta[0], ref_b = use_jd ? tb[0] : 0.0;
    double lo = 0.0, hi = 0.0;
    int ilo = 0, ihi = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double d = (ta[i] - ref_a) + ((use_jd ? tb[i] : 0.0) - ref_b);
        if (d < lo) { lo = d; ilo = i; }
        if (d > hi) { hi = d; ihi = i; }
    }
    s_lo[threadIdx.x] = lo; s_hi[threadIdx.x] = hi;
    s_ilo[threadIdx.x] = ilo; s_ihi[threadIdx.x] = ihi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            if (s_lo[threadIdx.x + w] < s_lo[threadIdx.x]) {
                s_lo[threadIdx.x] = s_lo[threadIdx.x + w];
                s_ilo[threadIdx.x] = s_ilo[threadIdx.x + w];
            }
            if (s_hi[threadIdx.x + w] > s_hi[threadIdx.x]) {
                s_hi[threadIdx.x] = s_hi[threadIdx.x + w];
                s_ihi[threadIdx.x] = s_ihi[threadIdx.x + w];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        range[0] = ta[s_ilo[0]];
        range[1] = use_jd ? tb[s_ilo[0]] : 0.0;
        range[2] = ta[s_ihi[0]];
        range[3] = use_jd ? tb[s_ihi[0]] : 0.0;
    }
}

// Pre-pass 2: one thread per resonant satellite walks the +-720 min integrator from epoch
// across the step indices the grid will need and tabulates (xli, xni) at each node, so that
// every (satellite, time) evaluation costs O(1) instead of |t|/720 integrator steps. The walk
// is inherently serial per satellite but short (14 nodes per week of grid).
__global__ void __launch_bounds__(128)
resonance_nodes_kernel(const double *__restrict__ pf, size_t n_pad, const int32_t *__restrict__ perm,
                       size_t first_slot, int n_slots, int first_r2_slot_rel,
                       const double *__restrict__ range, int use_jd, int cap,
                       ResNode *__restrict__ nodes, int *__restrict__ k0_out,
                       int *__restrict__ cnt_out) {
    const int rs = blockIdx.x * blockDim.x + threadIdx.x;
    if (rs >= n_slots) return;
    const size_t slot = first_slot + rs;
    if (perm[slot] < 0) {
        k0_out[rs] = 0;
        cnt_out[rs] = 0;
        return;
    }
    const GlobalConsts k = {pf + slot, n_pad};
    const int irez = (rs >= first_r2_slot_rel) ? 2 : 1;
    const double t_a = tsince_of(use_jd, range[0], range[1], k(PF_JD0), k(PF_JDF0));
    const double t_b = tsince_of(use_jd, range[2], range[3], k(PF_JD0), k(PF_JDF0));
    const int n_a = resonance_step_count(t_a), n_b = resonance_step_count(t_b);
    int k_lo = (t_a > 0.0) ? n_a : -n_a;
    int k_hi = (t_b > 0.0) ? n_b : -n_b;
    if (k_lo > k_hi) { const int tmp = k_lo; k_lo = k_hi; k_hi = tmp; }
    const int kMaxWalk = 1400000;  // resonance_step_count() never exceeds 1e9 / 720
    int count = k_hi - k_lo + 1;
    if (count > cap) count = cap;
    if (k_lo < -kMaxWalk || k_hi > kMaxWalk) count = 0;  // evaluations fall back to in-line steps
    k0_out[rs] = k_lo;
    cnt_out[rs] = count;
    if (count <= 0) return;
    const int k_end = k_lo + count;  // tabulate [k_lo, k_end)
    ResNode *mine = nodes + static_cast<size_t>(rs) * cap;
    const double no = k(PF_NO), xlamo = k(PF_XLAMO);
    // forwards from epoch
    if (k_end > 0) {
        double xli = xlamo, xni = no, atime = 0.0, xndt, xldot, xnddt;
        for (int kk = 0; kk < k_end; ++kk) {
            resonance_rates(k, irez, xli, xni, atime, xndt, xldot, xnddt);
            if (kk >= k_lo) mine[kk - k_lo] = {xli, xni, xndt, xnddt};
            resonance_advance(xndt, xldot, xnddt, 720.0, xli, xni, atime);
        }
    }
    // backwards from epoch
    if (k_lo < 0) {
        double xli = xlamo, xni = no, atime = 0.0, xndt, xldot, xnddt;
        for (int kk = 0; kk >= k_lo; --kk) {
            resonance_rates(k, irez, xli, xni, atime, xndt, xldot, xnddt);
            if (kk < k_end && kk <= 0 && !(kk == 0 && k_end > 0)) {
                mine[kk - k_lo] = {xli, xni, xndt, xnddt};
            }
            resonance_advance(xndt, xldot, xnddt, -720.0, xli, xni, atime);
        }
    }
}

// Output modes of kernel 2
constexpr int kModeStates = 0;    // six state planes + error plane
constexpr int kModeElements = 1;  // ... plus the seven mean-element planes
constexpr int kModeSummary = 2;   // fused consumer: per-satellite envelope, no [n][m] output
constexpr int kModeRecords = 3;   // array of 64-byte StateVector records + error plane
constexpr int kModeVisibility = 4;  // fused consumer: ground-station visibility bits, no [n][m] output
constexpr int kModeProximity = 5;   // fused consumer: closest approach to a few primary objects

// Warp-wide minimum of non-negative doubles (they order like their bit patterns) with ties
// going to the lowest index: three 32-bit reductions. Lanes without a candidate pass +inf, -1.
__device__ __forceinline__ void warp_min_with_index(double &v, int &k) {
    const unsigned hi = static_cast<unsigned>(__double2hiint(v));
    const unsigned lo = static_cast<unsigned>(__double2loint(v));
    const unsigned m_hi = warp_min_u32(hi);
    bool cand = hi == m_hi;
    const unsigned m_lo = warp_min_u32(cand ? lo : 0xffffffffu);
    cand = cand && lo == m_lo;
    const unsigned m_k = warp_min_u32(cand ? static_cast<unsigned>(k) : 0xffffffffu);
    v = __hiloint2double(static_cast<int>(m_hi), static_cast<int>(m_lo));
    k = static_cast<int>(m_k);
}

struct LaneSummary {
    double r_min, r_max;
    int k_min, k_max, n_ok, k_err, code_err;
};

// Combine two partial summaries; `b` covers later times than `a` only in the sense that ties
// go to the lower time index, which the explicit index comparisons implement.
__device__ __forceinline__ void merge_summary(LaneSummary &a, const LaneSummary &b) {
    if (b.r_min < a.r_min || (b.r_min == a.r_min && b.k_min >= 0 && (a.k_min < 0 || b.k_min < a.k_min))) {
        a.r_min = b.r_min;
        a.k_min = b.k_min;
    }
    if (b.r_max > a.r_max || (b.r_max == a.r_max && b.k_max >= 0 && (a.k_max < 0 || b.k_max < a.k_max))) {
        a.r_max = b.r_max;
        a.k_max = b.k_max;
    }
    a.n_ok += b.n_ok;
    if (b.k_err >= 0 && (a.k_err < 0 || b.k_err < a.k_err)) {
        a.k_err = b.k_err;
        a.code_err = b.code_err;
    }
}

// SPLIT: for a grid of at most 64 times (one time tile that it does not fill) the four warps
// of a CTA also split the tile's SATELLITES -- 4 (2) warps share a group of 32 times and take
// every 4th (2nd) satellite -- instead of leaving warps idle and the rest with all 32
// satellites to walk one after the other. A separate instantiation: the variable loop bounds
// cost the long-grid kernel registers (16 bytes of spills, 9 % slower when tried in place).
//
// SATS: satellites of the 32-satellite tile one CTA walks (32, 16 or 8). All CTAs of a launch
// take the same time, so the last, partly filled wave leaves SMs idle for up to one CTA
// duration: on a grid of few waves (30 000 satellites x 1440 times = 10.6 waves of 32-satellite
// CTAs) smaller CTAs shorten that tail for a slightly larger staging share.
template <int CLS, int MODE, bool SPLIT = false, int SATS = PB_TILE_SATS>
__global__ void __launch_bounds__(kThreads, (CLS <= PB_CLS_NEAR_SIMP) ? PB_MIN_BLOCKS_NEAR
                                           : (CLS == PB_CLS_DEEP_R2) ? PB_MIN_BLOCKS_R2
                                                                     : PB_MIN_BLOCKS_DEEP)
sgp4_prop_kernel(const PropParams p) {
    static_assert(SATS == 32 || SATS == 16 || SATS == 8, "CTA walks 32, 16 or 8 satellites");
    constexpr int kSub = PB_TILE_SATS / SATS;  // CTAs per (satellite tile, time tile)
    constexpr bool WANT_ELEMENTS = (MODE == kModeElements);
    using TF = TileFields<CLS>;
    constexpr int kRowsPad = (TF::kRows + 1) & ~1;  // even: every satellite's block 16 B aligned
    __shared__ __align__(16) double s_pf[kRowsPad * SATS];
    __shared__ double s_ta[kTimeTile];
    __shared__ double s_tb[kTimeTile];
    __shared__ int s_perm[SATS];
    constexpr bool kResonant = (CLS == PB_CLS_DEEP_R1 || CLS == PB_CLS_DEEP_R2);
    __shared__ int4 s_res[kResonant ? SATS : 1];  // {k0, count, first record, -}
    // deep space: sin / cos of the per-time and per-satellite parts of the solar and lunar mean
    // anomalies (ThirdBodyPhase, sgp4_eval.cuh): [0, N) Sun, [N, 2N) Moon
    constexpr bool kDeepCls = (CLS >= PB_CLS_DEEP_NR);
    __shared__ __align__(16) double2 s_tphase[kDeepCls ? 2 * kTimeTile : 1];
    __shared__ __align__(16) double2 s_sphase[kDeepCls ? 2 * SATS : 1];

    const int ttile = blockIdx.x % p.n_time_tiles;
    const int stile_sub = blockIdx.x / p.n_time_tiles;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const size_t sat0 = static_cast<size_t>(p.tile_begin) * PB_TILE_SATS +
                        static_cast<size_t>(stile_sub) * SATS;

    // ---- stage constants (coalesced 256 B rows) and this CTA's slice of the time grid ----
    // (a grid that fits one time tile is launched with just the warps it fills: blockDim.x
    // may be below kThreads, so the staging loops go by the actual CTA size)
    const int n_warps = static_cast<int>(blockDim.x >> 5);
    {   // a warp covers kSub rows of SATS satellites per pass
        const int sl = lane % SATS, rl = lane / SATS;
        for (int r = warp * kSub + rl; r < TF::kRows; r += n_warps * kSub) {
            const int f = r < TF::kEndA ? r : r - TF::kEndA + TF::kBeginB;
            s_pf[sl * kRowsPad + r] = p.pf[static_cast<size_t>(f) * p.n_pad + sat0 + sl];
        }
    }
    if (threadIdx.x < SATS) s_perm[threadIdx.x] = p.perm[sat0 + threadIdx.x];
    if (kResonant && threadIdx.x < SATS) {
        const size_t rs = sat0 + threadIdx.x - p.res_first_slot;
        const bool have = p.res_nodes != nullptr;
        s_res[threadIdx.x] = make_int4(have ? p.res_k0[rs] : 0, have ? p.res_cnt[rs] : 0,
                                       static_cast<int>(rs) * p.res_cap, 0);
    }
    const int k_base = ttile * kTimeTile;
    for (int i = threadIdx.x; i < static_cast<int>(blockDim.x) * kTpl; i += blockDim.x) {
        // Lanes past the end of the grid re-evaluate the last time (never stored): a made-up
        // time such as 0 would be ~1e9 minutes from epoch and spin the resonance integrator.
        const int k = min(k_base + i, p.m - 1);
        const double ta_k = p.ta[k], tb_k = p.use_jd ? p.tb[k] : 0.0;
        s_ta[i] = ta_k;
        s_tb[i] = tb_k;
        if (kDeepCls) {
            phase_of_time(phase_minutes_of_time(p.use_jd, ta_k, tb_k), s_tphase[i],
                          s_tphase[kTimeTile + i]);
        }
    }
    if (kDeepCls && threadIdx.x < SATS) {  // (a CTA has at least 32 threads)
        const double *f = p.pf + sat0 + threadIdx.x;
        phase_of_satellite(f[static_cast<size_t>(PF_ZMOS) * p.n_pad], f[static_cast<size_t>(PF_ZMOL) * p.n_pad],
                           phase_minutes_of_epoch(p.use_jd, f[static_cast<size_t>(PF_JD0) * p.n_pad],
                                                  f[static_cast<size_t>(PF_JDF0) * p.n_pad]),
                           s_sphase[threadIdx.x], s_sphase[SATS + threadIdx.x]);
    }
    __syncthreads();

    // Lane l of warp w owns time w*32 + l of the tile (with PB_TPL = 2: w*64 + l and
    // w*64 + 32 + l): every evaluation of a warp covers 32 CONSECUTIVE times, so its stores are
    // 256 contiguous bytes per plane (full sectors) issued right after the evaluation.
    int tgroup = warp, sat_first = 0, sat_step = 1;
    if (SPLIT) {  // launched only with kTpl == 1, one time tile and m <= 64
        const int groups = (p.m <= 32) ? 1 : 2;
        tgroup = warp % groups;
        sat_first = warp / groups;
        sat_step = kWarps / groups;
    }
    const int i0 = tgroup * (32 * kTpl) + (kPair ? 2 * lane : lane);
    constexpr int kTimeStep = kPair ? 1 : 32;  // distance between a lane's two times
    if (k_base + tgroup * 32 * kTpl >= p.m) return;  // whole warp beyond the grid
    // (With PB_TPL = 2, skipping the second group of 32 times when only it lies beyond the grid
    // was measured: a data-dependent trip count makes ptxas treat the body as divergent, +1.4 %
    // instructions per evaluation; one time per lane avoids the question.)
    double ta[kTpl], tb[kTpl];
#pragma unroll
    for (int j = 0; j < kTpl; ++j) {
        ta[j] = s_ta[i0 + kTimeStep * j];
        tb[j] = s_tb[i0 + kTimeStep * j];
    }
    const ClassFlags no_flags = {false, false, 0};
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double gm_s = 0.0, gm_c = 1.0;
    if (MODE == kModeVisibility) {
#ifndef PB_SWEEP_BUILD  // (kernel sweeps build PB_TPL = 2 libraries to time the states mode only)
        static_assert(MODE != kModeVisibility || kTpl == 1, "fused screens are built for one time per lane");
#endif
        const int kc = min(k_base + i0, p.m - 1);
        gm_s = p.gmst[2 * kc];
        gm_c = p.gmst[2 * kc + 1];
    }

#pragma unroll 1
    for (int s = SPLIT ? sat_first : 0; s < SATS; s += SPLIT ? sat_step : 1) {
        const int dst = s_perm[s];
        if (dst < 0) continue;  // padding slot (warp-uniform)
        // The block's offset is made opaque to the optimiser (an empty volatile asm): the
        // compiler then forms the block's shared-window address once per satellite and keeps it in
        // one uniform register across the evaluation; left to itself it re-formed it (S2UR
        // CgaCtaId, UMOV, ULEA, UIMAD) three times per evaluation -- 12 issue slots (SASS of
        // both forms compared with scripts/sass_path.py).
        int soff = s * kRowsPad;
#if PB_SMEM_VECTOR_ADDR
        asm volatile("" : "+r"(soff));
#endif
        const SmemConsts<CLS> k = {s_pf + soff};
        ResNodes nodes = {nullptr, 0, 0, 0};
        if (kResonant) {
            const int4 rn = s_res[s];
            nodes.table = p.res_nodes;
            nodes.k0 = rn.x;
            nodes.count = rn.y;
            nodes.first = rn.z;
        }
        constexpr bool kFused = (MODE == kModeSummary || MODE == kModeVisibility || MODE == kModeProximity);
        double *row = kFused                   ? nullptr
                      : (MODE == kModeRecords) ? p.states + static_cast<size_t>(dst) * p.row_stride * 8
                                               : p.out + static_cast<size_t>(dst) * p.row_stride;
        int32_t *erow = kFused ? nullptr : p.err + static_cast<size_t>(dst) * p.err_row_stride;
        LaneSummary acc = {__longlong_as_double(0x7ff0000000000000LL),
                           __longlong_as_double(0xfff0000000000000LL), -1, -1, 0, -1, 0};

        if (kPair && MODE == kModeStates) {
            // both evaluations in one instruction stream, then paired stores
            double r[2][6];
            int cd[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double t = tsince_of(p.use_jd, ta[j], tb[j], k(PF_JD0), k(PF_JDF0));
                ThirdBodyPhase ph;
                if (kDeepCls) {
                    ph = phase_combine(s_sphase[s], s_sphase[SATS + s], s_tphase[i0 + j],
                                       s_tphase[kTimeTile + i0 + j]);
                }
                cd[j] = sgp4_eval<CLS, false, true>(k, p.g, p.opsmode_a != 0, no_flags, t, r[j],
                                                    nullptr, kResonant ? &nodes : nullptr,
                                                    kDeepCls ? &ph : nullptr);
            }
            const int kt = k_base + i0;
            const bool w0 = (cd[0] == 0) || (cd[0] == 6), w1 = (cd[1] == 0) || (cd[1] == 6);
            if (!warp_all(w0 && w1)) {
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    r[0][c] = w0 ? r[0][c] : nan;
                    r[1][c] = w1 ? r[1][c] : nan;
                }
            }
            if (p.vec_ok && kt + 1 < p.m) {
                double *q = row + kt;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    *reinterpret_cast<double2 *>(q) = make_double2(r[0][c], r[1][c]);
                    q += p.plane_stride;
                }
                *reinterpret_cast<int2 *>(erow + kt) = make_int2(cd[0], cd[1]);
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (kt + j < p.m) {
#pragma unroll
                        for (int c = 0; c < 6; ++c) row[c * p.plane_stride + kt + j] = r[j][c];
                        erow[kt + j] = cd[j];
                    }
                }
            }
            continue;
        }
        // With PB_TPL = 2 (rolled) the two evaluations share ONE copy of the code: the body is
        // ~15 KB of SASS and the instruction cache is 32 KB.
#pragma unroll 1
        for (int j = 0; j < kTpl; ++j) {
            const double t = tsince_of(p.use_jd, (j == 0) ? ta[0] : ta[kTpl - 1], (j == 0) ? tb[0] : tb[kTpl - 1],
                                       k(PF_JD0), k(PF_JDF0));
            double r[6];
            Sgp4Side side;
            ThirdBodyPhase ph;
            if (kDeepCls) {
                ph = phase_combine(s_sphase[s], s_sphase[SATS + s], s_tphase[i0 + kTimeStep * j],
                                   s_tphase[kTimeTile + i0 + kTimeStep * j]);
            }
            const int cd = sgp4_eval<CLS, WANT_ELEMENTS, true>(
                k, p.g, p.opsmode_a != 0, no_flags, t, r, WANT_ELEMENTS ? &side : nullptr,
                kResonant ? &nodes : nullptr, kDeepCls ? &ph : nullptr);
            const int kt = k_base + i0 + kTimeStep * j;
            // every lane of the warp is here (uniform control flow): one vote decides whether
            // the stores need the NaN selects at all
            const bool all_written = warp_all((cd == 0) || (cd == 6));
            if (WANT_ELEMENTS && kt < p.m) {
                // what sgp4() leaves in satrec (src/sgp4.cpp:1895-1901); not reached for codes 1, 2
                const bool stored = side.stage >= 1;
                double *erow_el = p.elements + static_cast<size_t>(dst) * p.row_stride + kt;
                erow_el[0 * p.plane_stride] = stored ? side.am : nan;
                erow_el[1 * p.plane_stride] = stored ? side.em : nan;
                erow_el[2 * p.plane_stride] = stored ? side.im : nan;
                erow_el[3 * p.plane_stride] = stored ? side.Om : nan;
                erow_el[4 * p.plane_stride] = stored ? side.om : nan;
                erow_el[5 * p.plane_stride] = stored ? side.mm : nan;
                erow_el[6 * p.plane_stride] = stored ? side.nm : nan;
            }
            if (MODE == kModeVisibility) {
                // Earth-fixed position: TEME rotated about z by the sidereal angle of this
                // lane's time (sin / cos fetched once per CTA, the lane's time never changes)
                const bool ok = (cd == 0) && (kt < p.m);
                const double xe = mad2(gm_c, r[0], gm_s, r[1]);
                const double ye = msb2(gm_c, r[1], gm_s, r[0]);
                const int word = ttile * kWarps + tgroup;
                for (int st = 0; st < p.n_stations; ++st) {
                    const double *q = p.stations + 8 * st;
                    const double dx = xe - q[0], dy = ye - q[1], dz = r[2] - q[2];
                    const double range2 = __fma_rn(dz, dz, mad2(dx, dx, dy, dy));
                    const double up = __fma_rn(dz, q[5], mad2(dx, q[3], dy, q[4]));
                    // sin(elevation) = up / range against the mask, without the division
                    const double sin_el = up * rsqrt_fast(range2);
                    const bool visible = ok && sin_el >= q[6];
                    const unsigned bits = warp_ballot(visible);
                    // largest sin(elevation) of the 32 times: 1 - sin is non-negative
                    const double gap = 1.0 - sin_el;  // (rounding can leave sin_el a hair above 1)
                    double key = ok ? (gap > 0.0 ? gap : 0.0) : __longlong_as_double(0x7ff0000000000000LL);
                    int kbest = ok ? kt : -1;
                    warp_min_with_index(key, kbest);
                    if (lane == 0) {
                        const size_t at = (static_cast<size_t>(dst) * p.n_stations + st) * p.chunks + word;
                        p.vis_words[at] = bits;
                        p.vis_best[at] = make_double2(1.0 - key, static_cast<double>(kbest));
                    }
                }
            } else if (MODE == kModeProximity) {
                const bool ok = (cd == 0) && (kt < p.m);
                const int word = ttile * kWarps + tgroup;
                const int kc = kt < p.m ? kt : p.m - 1;
                for (int pr = 0; pr < p.n_primaries; ++pr) {
                    const double *q = p.primaries + static_cast<size_t>(pr) * 3 * p.m + kc;
                    const double dx = r[0] - q[0], dy = r[1] - q[p.m], dz = r[2] - q[2 * static_cast<size_t>(p.m)];
                    const double d2v = __fma_rn(dz, dz, mad2(dx, dx, dy, dy));
                    const bool have = ok && (d2v == d2v);  // a primary without a state is skipped
                    const unsigned below = warp_ballot(have && d2v < p.thr2);
                    double key = have ? d2v : __longlong_as_double(0x7ff0000000000000LL);
                    int kbest = have ? kt : -1;
                    warp_min_with_index(key, kbest);
                    if (lane == 0) {
                        ProxPartial out;
                        out.d2_min = key;
                        out.k_min = kbest;
                        out.n_below = __popc(below);
                        p.prox[(static_cast<size_t>(dst) * p.n_primaries + pr) * p.chunks + word] = out;
                    }
                }
            } else if (MODE == kModeSummary) {
                if (kt < p.m) {
                    if (cd == 0) {
                        const double rr = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                        if (rr < acc.r_min) { acc.r_min = rr; acc.k_min = kt; }
                        if (rr > acc.r_max) { acc.r_max = rr; acc.k_max = kt; }
                        acc.n_ok += 1;
                    } else if (acc.k_err < 0) {
                        acc.k_err = kt;  // j = 0 precedes j = 1 in time
                        acc.code_err = cd;
                    }
                }
            } else if (MODE == kModeRecords) {
                if (kt < p.m) {
                    // one perturb::StateVector per lane: 64 contiguous bytes, four 16-byte
                    // stores; a warp covers 2 KB of the satellite's row
                    const bool written = (cd == 0) || (cd == 6);
                    const double tj = (j == 0) ? ta[0] : ta[kTpl - 1];
                    double2 e;
                    if (p.use_jd) {
                        e.x = tj;  // sv.epoch = jd, perturb.cpp:240
                        e.y = (j == 0) ? tb[0] : tb[kTpl - 1];
                    } else {
                        e.x = k(PF_JD0);  // epoch() + mins / 1440, perturb.cpp:229 and :85-89
                        e.y = __dadd_rn(k(PF_JDF0), __ddiv_rn(tj, 1440.0));
                    }
                    double2 *rec = reinterpret_cast<double2 *>(row + static_cast<size_t>(kt) * 8);
                    rec[0] = e;
                    rec[1] = make_double2(written ? r[0] : nan, written ? r[1] : nan);
                    rec[2] = make_double2(written ? r[2] : nan, written ? r[3] : nan);
                    rec[3] = make_double2(written ? r[4] : nan, written ? r[5] : nan);
                    erow[kt] = cd;
                }
            } else if (kt < p.m) {
                // codes 1-4: the reference leaves the caller's state untouched -> NaN
                const bool written = (cd == 0) || (cd == 6);
                if (all_written) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) row[c * p.plane_stride + kt] = r[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 6; ++c) row[c * p.plane_stride + kt] = written ? r[c] : nan;
                }
                erow[kt] = cd;
            }
        }
        if (MODE == kModeSummary) {
            // Warp reduction over this warp's times of the satellite, then one partial record. The
            // radii are non-negative doubles, which order like their bit patterns, so min / max
            // with "ties to the lowest time index" are three 32-bit warp reductions each (high
            // word, low word among the lanes that match, index among those that match both)
            // instead of five rounds of seven shuffles and a full merge.
            {
                // minimum radius: lanes without an ok evaluation hold +inf and index -1
                unsigned hi = static_cast<unsigned>(__double2hiint(acc.r_min));
                unsigned lo = static_cast<unsigned>(__double2loint(acc.r_min));
                unsigned m_hi = warp_min_u32(hi);
                bool cand = hi == m_hi;
                unsigned m_lo = warp_min_u32(cand ? lo : 0xffffffffu);
                cand = cand && lo == m_lo;
                unsigned m_k = warp_min_u32(cand ? static_cast<unsigned>(acc.k_min) : 0xffffffffu);
                acc.r_min = __hiloint2double(static_cast<int>(m_hi), static_cast<int>(m_lo));
                acc.k_min = static_cast<int>(m_k);
                // maximum radius: lanes without an ok evaluation (-inf, index -1) count as key 0
                const bool have = acc.k_max >= 0;
                hi = have ? static_cast<unsigned>(__double2hiint(acc.r_max)) : 0u;
                lo = have ? static_cast<unsigned>(__double2loint(acc.r_max)) : 0u;
                m_hi = warp_max_u32(hi);
                cand = hi == m_hi;
                m_lo = warp_max_u32(cand ? lo : 0u);
                cand = cand && lo == m_lo;
                m_k = warp_min_u32(cand ? static_cast<unsigned>(acc.k_max) : 0xffffffffu);
                acc.k_max = static_cast<int>(m_k);
                acc.r_max = (acc.k_max >= 0)
                                ? __hiloint2double(static_cast<int>(m_hi), static_cast<int>(m_lo))
                                : __longlong_as_double(0xfff0000000000000LL);
                acc.n_ok = warp_add_s32(acc.n_ok);
                // first error in time: every lane's index is distinct, so one lane matches
                const unsigned ke = static_cast<unsigned>(acc.k_err);  // -1 -> 0xffffffff
                const unsigned m_e = warp_min_u32(ke);
                const unsigned code = warp_max_u32(
                    (ke == m_e && acc.k_err >= 0) ? static_cast<unsigned>(acc.code_err) : 0u);
                acc.k_err = static_cast<int>(m_e);
                acc.code_err = static_cast<int>(code);
            }
            if (lane == 0) {
                perturb_gpu_summary out;
                out.r_min = acc.r_min;
                out.r_max = acc.r_max;
                out.k_min = acc.k_min;
                out.k_max = acc.k_max;
                out.n_ok = acc.n_ok;
                out.first_error = acc.code_err;
                out.k_first_error = acc.k_err;
                out.reserved = 0;
                p.partials[static_cast<size_t>(dst) * p.chunks + (ttile * kWarps + tgroup)] = out;
            }
        }
    }
}

// Kernel 2 for SHORT time grids (a whole catalogue at one epoch, or a handful): with lanes =
// times a warp of the kernel above would have m of its 32 lanes busy. Here lanes are 32
// class-sorted SATELLITES of one tile and every lane walks the few times; the constants come
// straight from the field-major layout (one coalesced 256 B row per field, L1 resident across
// the times). The evaluation is the same code on the same inputs, so the result of a
// (satellite, time) pair is bit-identical to the other kernel's (tested).
#ifndef PB_SHORT_GRID
#define PB_SHORT_GRID 13
#endif
constexpr int kShortGrid = PB_SHORT_GRID;  // grids up to this many times take this kernel (crossover
                                          // with the SPLIT lanes = times kernel measured at 13-14)

template <int CLS, int MODE>
__global__ void __launch_bounds__(kThreads)
sgp4_prop_short_kernel(const PropParams p, int tile_count) {
    constexpr bool WANT_ELEMENTS = (MODE == kModeElements);
    constexpr bool kResonant = (CLS == PB_CLS_DEEP_R1 || CLS == PB_CLS_DEEP_R2);
    constexpr bool kDeepCls = (CLS >= PB_CLS_DEEP_NR);
    // the same per-time / per-satellite split of the solar and lunar mean anomalies as the
    // long-grid kernel (same functions on the same inputs: same bits)
    __shared__ __align__(16) double2 s_tphase[kDeepCls ? 2 * kShortGrid : 1];
    if (kDeepCls) {
        if (static_cast<int>(threadIdx.x) < p.m) {
            const int kt = threadIdx.x;
            phase_of_time(phase_minutes_of_time(p.use_jd, p.ta[kt], p.use_jd ? p.tb[kt] : 0.0),
                          s_tphase[kt], s_tphase[kShortGrid + kt]);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (warp_all(tile >= tile_count)) return;
    const size_t slot_own = static_cast<size_t>(p.tile_begin + tile) * PB_TILE_SATS + lane;
    const int dst = p.perm[slot_own];
    // a padding lane (only at the end of a class segment, never first in its tile) redoes the
    // tile's first satellite and stores nothing: every lane stays active for the votes
    const size_t slot = (dst >= 0) ? slot_own : slot_own - lane;
    const GlobalConsts k = {p.pf + slot, p.n_pad};
    ResNodes nodes = {nullptr, 0, 0, 0};
    if (kResonant && p.res_nodes != nullptr) {
        const size_t rs = slot - p.res_first_slot;
        nodes.table = p.res_nodes;
        nodes.k0 = p.res_k0[rs];
        nodes.count = p.res_cnt[rs];
        nodes.first = static_cast<int>(rs) * p.res_cap;
    }
    const ClassFlags no_flags = {false, false, 0};
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double jd0 = k(PF_JD0), jdf0 = k(PF_JDF0);
    double2 sat_sun = make_double2(0.0, 1.0), sat_moon = make_double2(0.0, 1.0);
    if (kDeepCls) {
        phase_of_satellite(k(PF_ZMOS), k(PF_ZMOL), phase_minutes_of_epoch(p.use_jd, jd0, jdf0),
                           sat_sun, sat_moon);
    }
    LaneSummary acc = {__longlong_as_double(0x7ff0000000000000LL),
                       __longlong_as_double(0xfff0000000000000LL), -1, -1, 0, -1, 0};
#pragma unroll 1
    for (int kt = 0; kt < p.m; ++kt) {
        const double ta = p.ta[kt], tb = p.use_jd ? p.tb[kt] : 0.0;
        const double t = tsince_of(p.use_jd, ta, tb, jd0, jdf0);
        double r[6];
        Sgp4Side side;
        ThirdBodyPhase ph;
        if (kDeepCls) ph = phase_combine(sat_sun, sat_moon, s_tphase[kt], s_tphase[kShortGrid + kt]);
        const int cd = sgp4_eval<CLS, WANT_ELEMENTS, true>(
            k, p.g, p.opsmode_a != 0, no_flags, t, r, WANT_ELEMENTS ? &side : nullptr,
            (kResonant && nodes.table != nullptr) ? &nodes : nullptr, kDeepCls ? &ph : nullptr);
        if (dst < 0) continue;
        const bool written = (cd == 0) || (cd == 6);
        if (MODE == kModeSummary) {
            if (cd == 0) {
                const double rr = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                if (rr < acc.r_min) { acc.r_min = rr; acc.k_min = kt; }
                if (rr > acc.r_max) { acc.r_max = rr; acc.k_max = kt; }
                acc.n_ok += 1;
            } else if (acc.k_err < 0) {
                acc.k_err = kt;
                acc.code_err = cd;
            }
        } else if (MODE == kModeRecords) {
            double2 e;
            if (p.use_jd) {
                e.x = ta;  // sv.epoch = jd, perturb.cpp:240
                e.y = tb;
            } else {
                e.x = jd0;  // epoch() + mins / 1440, perturb.cpp:229 and :85-89
                e.y = __dadd_rn(jdf0, __ddiv_rn(ta, 1440.0));
            }
            double2 *rec = reinterpret_cast<double2 *>(
                p.states + (static_cast<size_t>(dst) * p.row_stride + kt) * 8);
            rec[0] = e;
            rec[1] = make_double2(written ? r[0] : nan, written ? r[1] : nan);
            rec[2] = make_double2(written ? r[2] : nan, written ? r[3] : nan);
            rec[3] = make_double2(written ? r[4] : nan, written ? r[5] : nan);
            p.err[static_cast<size_t>(dst) * p.err_row_stride + kt] = cd;
        } else {
            double *row = p.out + static_cast<size_t>(dst) * p.row_stride + kt;
#pragma unroll
            for (int c = 0; c < 6; ++c) row[c * p.plane_stride] = written ? r[c] : nan;
            p.err[static_cast<size_t>(dst) * p.err_row_stride + kt] = cd;
            if (WANT_ELEMENTS) {
                const bool stored = side.stage >= 1;
                double *el = p.elements + static_cast<size_t>(dst) * p.row_stride + kt;
                el[0 * p.plane_stride] = stored ? side.am : nan;
                el[1 * p.plane_stride] = stored ? side.em : nan;
                el[2 * p.plane_stride] = stored ? side.im : nan;
                el[3 * p.plane_stride] = stored ? side.Om : nan;
                el[4 * p.plane_stride] = stored ? side.om : nan;
                el[5 * p.plane_stride] = stored ? side.mm : nan;
                el[6 * p.plane_stride] = stored ? side.nm : nan;
            }
        }
    }
    if (MODE == kModeSummary && dst >= 0) {
        // the whole (short) grid is this lane's: chunk 0 holds the satellite's summary
        perturb_gpu_summary out;
        out.r_min = acc.r_min;
        out.r_max = acc.r_max;
        out.k_min = acc.k_min;
        out.k_max = acc.k_max;
        out.n_ok = acc.n_ok;
        out.first_error = acc.code_err;
        out.k_first_error = acc.k_err;
        out.reserved = 0;
        p.partials[static_cast<size_t>(dst) * p.chunks] = out;
    }
}

// Second pass of the summary mode: one thread per satellite folds its chunk partials.
__global__ void __launch_bounds__(128)
summary_reduce_kernel(const perturb_gpu_summary *__restrict__ partials, int chunks, int used_chunks,
                      size_t n, perturb_gpu_summary *__restrict__ out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    LaneSummary acc = {__longlong_as_double(0x7ff0000000000000LL),
                       __longlong_as_double(0xfff0000000000000LL), -1, -1, 0, -1, 0};
    for (int c = 0; c < used_chunks; ++c) {
        const perturb_gpu_summary q = partials[i * chunks + c];
        const LaneSummary o = {q.r_min, q.r_max, q.k_min, q.k_max, q.n_ok, q.k_first_error,
                               q.first_error};
        merge_summary(acc, o);
    }
    perturb_gpu_summary r;
    r.r_min = acc.r_min;
    r.r_max = acc.r_max;
    r.k_min = acc.k_min;
    r.k_max = acc.k_max;
    r.n_ok = acc.n_ok;
    r.first_error = acc.code_err;
    r.k_first_error = acc.k_err;
    r.reserved = 0;
    out[i] = r;
}

// Second pass of the visibility screen: one thread per (satellite, station) walks the visibility
// bits in time order -- runs of ones are passes -- and folds the per-chunk elevation maxima.
__global__ void __launch_bounds__(128)
visibility_reduce_kernel(const uint32_t *__restrict__ words, const double2 *__restrict__ best,
                         int chunks, int used_chunks, size_t pairs, int m,
                         perturb_gpu_pass_summary *__restrict__ out, int32_t *__restrict__ passes,
                         int max_passes) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= pairs) return;
    const uint32_t *w = words + i * chunks;
    const double2 *b = best + i * chunks;
    int32_t *mine = passes != nullptr ? passes + i * static_cast<size_t>(max_passes) * 2 : nullptr;
    perturb_gpu_pass_summary r;
    r.max_sin_elevation = __longlong_as_double(0xfff0000000000000LL);
    r.k_max_elevation = -1;
    r.n_visible = 0;
    r.n_passes = 0;
    r.k_first_rise = -1;
    r.k_last_visible = -1;
    r.reserved = 0;
    bool in_pass = false;
    for (int c = 0; c < used_chunks; ++c) {
        const double2 q = b[c];
        if (q.y >= 0.0 && q.x > r.max_sin_elevation) {  // strict: ties stay with the earlier chunk
            r.max_sin_elevation = q.x;
            r.k_max_elevation = static_cast<int>(q.y);
        }
        const uint32_t bits = w[c];
        r.n_visible += __popc(bits);
        if (bits != 0u) r.k_last_visible = c * 32 + 31 - __clz(bits);
        int pos = 0;
        while (pos < 32) {
            const uint32_t rest = bits >> pos;
            if (!in_pass) {
                if (rest == 0u) break;
                pos += __ffs(rest) - 1;  // a rise
                if (r.k_first_rise < 0) r.k_first_rise = c * 32 + pos;
                if (mine != nullptr && r.n_passes < max_passes) mine[2 * r.n_passes] = c * 32 + pos;
                r.n_passes += 1;
                in_pass = true;
            } else {
                int run = __ffs(~rest) - 1;  // ones from `pos` on (zeros were shifted in on top)
                if (run < 0) run = 32;
                if (pos + run >= 32) break;  // the pass continues in the next chunk
                pos += run;                   // first invisible time: the pass ended one before
                if (mine != nullptr && r.n_passes <= max_passes) mine[2 * r.n_passes - 1] = c * 32 + pos - 1;
                in_pass = false;
            }
        }
    }
    if (in_pass && mine != nullptr && r.n_passes <= max_passes) {
        mine[2 * r.n_passes - 1] = m - 1;  // still above the mask at the end of the grid
    }
    if (mine != nullptr) {
        for (int j = r.n_passes; j < max_passes; ++j) {
            mine[2 * j] = -1;
            mine[2 * j + 1] = -1;
        }
    }
    out[i] = r;
}

// Second pass of the proximity screen: one thread per (satellite, primary) folds its chunk
// partials in time order (ties to the lowest index).
__global__ void __launch_bounds__(128)
proximity_reduce_kernel(const ProxPartial *__restrict__ partials, int chunks, int used_chunks,
                        size_t pairs, perturb_gpu_approach *__restrict__ out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= pairs) return;
    double d2 = __longlong_as_double(0x7ff0000000000000LL);
    int k = -1, below = 0;
    for (int c = 0; c < used_chunks; ++c) {
        const ProxPartial q = partials[i * chunks + c];
        if (q.k_min >= 0 && q.d2_min < d2) {
            d2 = q.d2_min;
            k = q.k_min;
        }
        below += q.n_below;
    }
    perturb_gpu_approach r;
    r.d_min_km = sqrt(d2);
    r.k_min = k;
    r.n_below = below;
    out[i] = r;
}

// Write-back pass: what ONE `sat.propagate(t, sv)` call leaves in
// `Satellite::sat_rec` besides the state -- t, error, the singly averaged mean elements
// (src/sgp4.cpp:1801-1802, 1895-1901), for deep-space satellites aycof / xlcof / con41 / x1mth2 /
// x7thm1 (:1952-1957, 2017-2019) and for resonant ones the integrator state atime / xli / xni
// (:1143-1147). One thread per class-sorted slot, class flags at run time; planes in caller
// order: wb[WB_COUNT][n].
__global__ void __launch_bounds__(128)
write_back_kernel(const double *__restrict__ pf, size_t n_pad, const int32_t *__restrict__ perm,
                  const uint8_t *__restrict__ cls, size_t n, double ta, double tb, int use_jd,
                  GravConsts g, int opsmode_a, double *__restrict__ wb) {
    const size_t slot = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (slot >= n_pad) return;
    const int dst = perm[slot];
    if (dst < 0) return;
    const GlobalConsts k = {pf + slot, n_pad};
    const int c = cls[dst];
    const ClassFlags rt = {c >= PB_CLS_DEEP_NR, c != PB_CLS_NEAR_FULL,
                           c == PB_CLS_DEEP_R1 ? 1 : c == PB_CLS_DEEP_R2 ? 2 : 0};
    const double t = tsince_of(use_jd, ta, tb, k(PF_JD0), k(PF_JDF0));
    double r[6];
    Sgp4Side side;
    const int code = sgp4_eval<PB_CLS_RUNTIME, true, false>(k, g, opsmode_a != 0, rt, t, r, &side);
    double *o = wb + dst;
    o[WB_T * n] = t;
    o[WB_ERROR * n] = static_cast<double>(code);
    o[WB_STAGE * n] = static_cast<double>(side.stage);
    o[WB_AM * n] = side.am;
    o[WB_EM * n] = side.em;
    o[WB_IM * n] = side.im;
    o[WB_OMEGA * n] = side.Om;
    o[WB_ARGP * n] = side.om;
    o[WB_MM * n] = side.mm;
    o[WB_NM * n] = side.nm;
    o[WB_AYCOF * n] = side.aycof;
    o[WB_XLCOF * n] = side.xlcof;
    o[WB_CON41 * n] = side.con41;
    o[WB_X1MTH2 * n] = side.x1mth2;
    o[WB_X7THM1 * n] = side.x7thm1;
    o[WB_RESONANT * n] = side.resonant ? 1.0 : 0.0;
    o[WB_ATIME * n] = side.atime;
    o[WB_XLI * n] = side.xli;
    o[WB_XNI * n] = side.xni;
}

template <int CLS>
void launch_class(const PropLaunch &l, PropParams p, cudaStream_t main, int &launches) {
    const int tiles = l.tile_count[CLS];
    if (tiles <= 0) return;
    p.tile_begin = l.tile_begin[CLS];
    const long long blocks = static_cast<long long>(tiles) * p.n_time_tiles;
    // grid.x limit is 2^31-1; at 1M satellites x 10k times this is 31250 * 40 blocks.
    cudaStream_t stream = main;
    const bool forked = (launches > 0) && l.aux[CLS] != nullptr && l.fork != nullptr &&
                        l.fork_recorded;
    if (forked) {
        stream = l.aux[CLS];
        if (cudaStreamWaitEvent(stream, l.fork, 0) != cudaSuccess) {
            stream = main;  // could not fork: run this class in line
        }
    }
    if (p.vis_words != nullptr || p.prox != nullptr) {
        // fused screens: lanes = times, whole tiles, only the warps a single time tile fills
        const unsigned threads = (p.n_time_tiles == 1)
                                     ? 32u * static_cast<unsigned>((p.m + 31) / 32)
                                     : static_cast<unsigned>(kThreads);
        const unsigned nb = static_cast<unsigned>(blocks);
        if (p.vis_words != nullptr) {
            sgp4_prop_kernel<CLS, kModeVisibility><<<nb, threads, 0, stream>>>(p);
        } else {
            sgp4_prop_kernel<CLS, kModeProximity><<<nb, threads, 0, stream>>>(p);
        }
    } else if (p.m <= kShortGrid) {
        const unsigned sb = static_cast<unsigned>((tiles + kWarps - 1) / kWarps);
        if (p.partials != nullptr) {
            sgp4_prop_short_kernel<CLS, kModeSummary><<<sb, kThreads, 0, stream>>>(p, tiles);
        } else if (p.states != nullptr) {
            sgp4_prop_short_kernel<CLS, kModeRecords><<<sb, kThreads, 0, stream>>>(p, tiles);
        } else if (p.elements != nullptr) {
            sgp4_prop_short_kernel<CLS, kModeElements><<<sb, kThreads, 0, stream>>>(p, tiles);
        } else {
            sgp4_prop_short_kernel<CLS, kModeStates><<<sb, kThreads, 0, stream>>>(p, tiles);
        }
    } else {
        // one time tile that the grid does not fill: CTAs of only the warps with work, so an SM
        // holds 4x / 2x as many satellite tiles (32 times: 28 busy warps per SM instead of 7)
        unsigned threads = kThreads;
        if (p.n_time_tiles == 1) {
            threads = 32u * static_cast<unsigned>((p.m + 32 * kTpl - 1) / (32 * kTpl));
        }
        const unsigned nb = static_cast<unsigned>(blocks);
        const bool split = (kTpl == 1) && (kWarps == 4) && (p.n_time_tiles == 1) && (p.m <= 64);
        if (split) {
            if (p.partials != nullptr) {
                sgp4_prop_kernel<CLS, kModeSummary, true><<<nb, kThreads, 0, stream>>>(p);
            } else if (p.states != nullptr) {
                sgp4_prop_kernel<CLS, kModeRecords, true><<<nb, kThreads, 0, stream>>>(p);
            } else if (p.elements != nullptr) {
                sgp4_prop_kernel<CLS, kModeElements, true><<<nb, kThreads, 0, stream>>>(p);
            } else {
                sgp4_prop_kernel<CLS, kModeStates, true><<<nb, kThreads, 0, stream>>>(p);
            }
        } else {
            // CTA size in satellites: 32 when the launch is many waves long anyway, else 16 / 8
            int sats = l.cta_sats;
            if (sats != 32 && sats != 16 && sats != 8) {
                const long long per_wave = 148LL * 6;  // SMs x resident CTAs, roughly
                sats = (blocks >= 40 * per_wave) ? 32 : (blocks >= 20 * per_wave) ? 16 : 8;
            }
            if (p.n_time_tiles == 1 || blocks * 4 > 0x7fffffffLL) sats = 32;
            const unsigned nbs = static_cast<unsigned>(blocks * (PB_TILE_SATS / sats));
#define PB_LAUNCH_MODE(MODE_)                                                                   \
    do {                                                                                        \
        if (sats == 32) {                                                                       \
            sgp4_prop_kernel<CLS, MODE_, false, 32><<<nbs, threads, 0, stream>>>(p);            \
        } else if (sats == 16) {                                                                \
            sgp4_prop_kernel<CLS, MODE_, false, 16><<<nbs, threads, 0, stream>>>(p);            \
        } else {                                                                                \
            sgp4_prop_kernel<CLS, MODE_, false, 8><<<nbs, threads, 0, stream>>>(p);             \
        }                                                                                       \
    } while (0)
            if (p.partials != nullptr) {
                PB_LAUNCH_MODE(kModeSummary);
            } else if (p.states != nullptr) {
                PB_LAUNCH_MODE(kModeRecords);
            } else if (p.elements != nullptr) {
                PB_LAUNCH_MODE(kModeElements);
            } else {
                PB_LAUNCH_MODE(kModeStates);
            }
#undef PB_LAUNCH_MODE
        }
    }
    if (stream != main) {
        // a failure here is sticky and surfaces in the caller's cudaGetLastError()
        if (cudaEventRecord(l.join[CLS], stream) == cudaSuccess) {
            cudaStreamWaitEvent(main, l.join[CLS], 0);
        }
    }
    ++launches;
}

}  // namespace

void launch_write_back(const PropLaunch &l, const uint8_t *d_cls, double ta, double tb,
                       double *d_wb, cudaStream_t stream) {
    if (l.n == 0 || l.n_pad == 0) return;
    GravConsts g;
    g.radiusearthkm = l.radiusearthkm;
    g.xke = l.xke;
    g.j2 = l.j2;
    g.j3oj2 = l.j3oj2;
    g.vkmpersec = l.radiusearthkm * l.xke / 60.0;
    g.inv_xke = 1.0 / l.xke;
    write_back_kernel<<<static_cast<unsigned>((l.n_pad + 127) / 128), 128, 0, stream>>>(
        l.pf, l.n_pad, l.perm, d_cls, l.n, ta, tb, l.use_jd, g, l.opsmode_a, d_wb);
}

size_t prox_partial_bytes() { return sizeof(ProxPartial); }

size_t summary_chunks(size_t m) {
    return ((m + kTimeTile - 1) / kTimeTile) * kWarps;
}

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
    p.elements = l.elements;
    p.states = l.states;
    p.partials = static_cast<perturb_gpu_summary *>(l.summary_partials);
    p.chunks = static_cast<int>(summary_chunks(static_cast<size_t>(l.m)));
    p.stations = l.stations;
    p.n_stations = l.n_stations;
    p.gmst = l.gmst;
    p.vis_words = l.vis_words;
    p.vis_best = l.vis_best;
    p.primaries = l.primaries;
    p.n_primaries = l.n_primaries;
    p.thr2 = l.threshold_km * l.threshold_km;
    p.prox = static_cast<ProxPartial *>(l.prox_partials);
    p.g.radiusearthkm = l.radiusearthkm;
    p.g.xke = l.xke;
    p.g.j2 = l.j2;
    p.g.j3oj2 = l.j3oj2;
    p.g.vkmpersec = l.radiusearthkm * l.xke / 60.0;
    p.g.inv_xke = 1.0 / l.xke;
    p.opsmode_a = l.opsmode_a;
    p.vec_ok = (l.out != nullptr && (reinterpret_cast<uintptr_t>(l.out) & 15u) == 0 &&
                (reinterpret_cast<uintptr_t>(l.err) & 7u) == 0 && l.row_stride % 2 == 0 &&
                l.plane_stride % 2 == 0 && l.err_row_stride % 2 == 0)
                   ? 1
                   : 0;
    int launches = 0;
    p.res_nodes = nullptr;
    p.res_k0 = p.res_cnt = nullptr;
    p.res_cap = 0;
    p.res_first_slot = 0;
    const int res_tiles = l.tile_count[PB_CLS_DEEP_R1] + l.tile_count[PB_CLS_DEEP_R2];
    if (res_tiles > 0 && l.res_nodes != nullptr && l.res_cap > 0) {
        const int n_slots = res_tiles * PB_TILE_SATS;
        const size_t first_slot = static_cast<size_t>(l.tile_begin[PB_CLS_DEEP_R1]) * PB_TILE_SATS;
        time_range_kernel<<<1, 256, 0, stream>>>(l.ta, l.tb, l.m, l.use_jd, l.res_range);
        resonance_nodes_kernel<<<(n_slots + 127) / 128, 128, 0, stream>>>(
            l.pf, l.n_pad, l.perm, first_slot, n_slots,
            l.tile_count[PB_CLS_DEEP_R1] * PB_TILE_SATS, l.res_range, l.use_jd, l.res_cap,
            reinterpret_cast<ResNode *>(l.res_nodes), l.res_k0, l.res_cnt);
        launches += 2;
        p.res_nodes = reinterpret_cast<const ResNode *>(l.res_nodes);
        p.res_cap = l.res_cap;
        p.res_k0 = l.res_k0;
        p.res_cnt = l.res_cnt;
        p.res_first_slot = first_slot;
    }
    const int prelaunches = launches;
    launches = 0;
    // grid.x is limited to 2^31 - 1 CTAs (a CTA may cover as few as 8 satellites x 128 times)
    for (int c = 0; c < PB_CLS_COUNT; ++c) {
        if (static_cast<long long>(l.tile_count[c]) * p.n_time_tiles > 0x7fffffffLL) return -1;
    }
    PropLaunch lf = l;
    lf.fork_recorded = (l.fork != nullptr) && cudaEventRecord(l.fork, stream) == cudaSuccess;
    // most expensive per evaluation first
    launch_class<PB_CLS_DEEP_R2>(lf, p, stream, launches);
    launch_class<PB_CLS_DEEP_R1>(lf, p, stream, launches);
    launch_class<PB_CLS_DEEP_NR>(lf, p, stream, launches);
    launch_class<PB_CLS_NEAR_FULL>(lf, p, stream, launches);
    launch_class<PB_CLS_NEAR_SIMP>(lf, p, stream, launches);
    const int used_chunks = static_cast<int>((static_cast<size_t>(l.m) + 31) / 32);
    if (l.vis_words != nullptr && l.pass_out != nullptr && l.n > 0) {
        const size_t pairs = l.n * static_cast<size_t>(l.n_stations);
        visibility_reduce_kernel<<<static_cast<unsigned>((pairs + 127) / 128), 128, 0, stream>>>(
            l.vis_words, l.vis_best, p.chunks, used_chunks, pairs, l.m,
            static_cast<perturb_gpu_pass_summary *>(l.pass_out), l.passes, l.max_passes);
        ++launches;
    }
    if (l.prox_partials != nullptr && l.approach_out != nullptr && l.n > 0) {
        const size_t pairs = l.n * static_cast<size_t>(l.n_primaries);
        proximity_reduce_kernel<<<static_cast<unsigned>((pairs + 127) / 128), 128, 0, stream>>>(
            p.prox, p.chunks, used_chunks, pairs, static_cast<perturb_gpu_approach *>(l.approach_out));
        ++launches;
    }
    if (l.summary_partials != nullptr && l.summary_out != nullptr && l.n > 0) {
        // every (satellite, chunk) partial with chunk * 32 * kTpl < m has been written by exactly one
        // warp of exactly one class launch; all of them were joined back into `stream`
        const int used = static_cast<int>((static_cast<size_t>(l.m) + 32 * kTpl - 1) / (32 * kTpl));
        summary_reduce_kernel<<<static_cast<unsigned>((l.n + 127) / 128), 128, 0, stream>>>(
            p.partials, p.chunks, used, l.n, static_cast<perturb_gpu_summary *>(l.summary_out));
        ++launches;
    }
    return launches + prelaunches;
}

}  // namespace pb
