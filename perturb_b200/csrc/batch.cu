// Host side of the C ABI (include/perturb_gpu.h): handle management, classification + sort,
// launches, host<->device staging. C++11, no exceptions across the boundary.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/perturb_gpu.h"
#include "../../include/perturb_elsetrec_layout.h"
#include "batch_internal.h"
#include "host_workers.h"

namespace {

thread_local std::string g_last_error = "";

int fail(int status, const std::string &msg) {
    g_last_error = msg;
    return status;
}

}  // namespace

int pb::report_failure(int status, const char *msg) { return fail(status, msg ? msg : ""); }

namespace {

#define PB_CUDA(call)                                                                        \
    do {                                                                                     \
        const cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                             \
            return fail(PERTURB_GPU_ERR_CUDA,                                                \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                 \
        }                                                                                    \
    } while (0)

const char *const kRecordFieldNames[RF_COUNT] = {
#define X(name) #name,
    PB_RECORD_FIELDS(X)
#undef X
};

bool is_device_pointer(const void *p) {
    if (p == nullptr) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// memcpy for the scatter of bounce chunks into the caller's arrays: the destination is written
// once and not read again soon, so 16-byte non-temporal stores skip the read-for-ownership a
// cached store costs -- on a box whose host memory bandwidth is what bounds the pipeline (DMA
// write + this copy's read and write) that is a quarter of the traffic. Falls back to memcpy for
// destinations that are not 16-byte aligned and for the tail.
inline void copy_streaming(void *dst, const void *src, size_t bytes) {
#if defined(__SSE2__)
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const size_t blocks = bytes / 64;
        const __m128i *s = static_cast<const __m128i *>(src);
        __m128i *d = static_cast<__m128i *>(dst);
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_load_si128(s + 4 * i), b = _mm_load_si128(s + 4 * i + 1);
            const __m128i c = _mm_load_si128(s + 4 * i + 2), e = _mm_load_si128(s + 4 * i + 3);
            _mm_stream_si128(d + 4 * i, a);
            _mm_stream_si128(d + 4 * i + 1, b);
            _mm_stream_si128(d + 4 * i + 2, c);
            _mm_stream_si128(d + 4 * i + 3, e);
        }
        const size_t done = blocks * 64;
        if (done < bytes) {
            std::memcpy(static_cast<char *>(dst) + done, static_cast<const char *>(src) + done, bytes - done);
        }
        return;
    }
#endif
    std::memcpy(dst, src, bytes);
}

inline void streaming_fence() {
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

// Page-locked host memory (cudaHostAlloc / cudaHostRegister): the copy engines reach it directly.
// Anything else on the host (malloc, std::vector, numpy) is pageable.
bool is_pinned_host_pointer(const void *p) {
    if (p == nullptr) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// Ordinal of the device whose memory `p` points into; -1 for host (or managed) memory.
int pointer_device(const void *p) {
    if (p == nullptr) return -1;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return attr.type == cudaMemoryTypeDevice ? attr.device : -1;
}

// A device buffer the kernels of `device` can read in place: its own memory or managed memory.
bool is_local_device_pointer(const void *p, int device) {
    if (!is_device_pointer(p)) return false;
    const int owner = pointer_device(p);
    return owner < 0 || owner == device;
}

// Kernels of `device` (the current device) are about to store into a buffer that lives on
// `owner`: a shard writing its satellite rows straight into the gathered result on another
// GPU of the same process, over NVLink peer access. Enabled on first use.
int ensure_peer_access(int device, int owner) {
    if (owner < 0 || owner == device) return PERTURB_GPU_OK;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, owner) != cudaSuccess || !can) {
        cudaGetLastError();
        return fail(PERTURB_GPU_ERR_ARG, "output buffer lives on device " + std::to_string(owner) +
                                             ", which device " + std::to_string(device) +
                                             " cannot access as a peer");
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        return fail(PERTURB_GPU_ERR_CUDA,
                    std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    }
    cudaGetLastError();
    return PERTURB_GPU_OK;
}

// getgravconst, src/sgp4.cpp:2101-2157
bool gravity_constants(int model, pb::GravAll &g) {
    switch (model) {
        case PERTURB_GPU_WGS72_OLD:
            g.mus = 398600.79964;
            g.radiusearthkm = 6378.135;
            g.xke = 0.0743669161;
            g.j2 = 0.001082616;
            g.j3 = -0.00000253881;
            g.j4 = -0.00000165597;
            break;
        case PERTURB_GPU_WGS72:
            g.mus = 398600.8;
            g.radiusearthkm = 6378.135;
            g.xke = 60.0 / std::sqrt(g.radiusearthkm * g.radiusearthkm * g.radiusearthkm / g.mus);
            g.j2 = 0.001082616;
            g.j3 = -0.00000253881;
            g.j4 = -0.00000165597;
            break;
        case PERTURB_GPU_WGS84:
            g.mus = 398600.5;
            g.radiusearthkm = 6378.137;
            g.xke = 60.0 / std::sqrt(g.radiusearthkm * g.radiusearthkm * g.radiusearthkm / g.mus);
            g.j2 = 0.00108262998905;
            g.j3 = -0.00000253215306;
            g.j4 = -0.00000161098761;
            break;
        default: return false;
    }
    g.tumin = 1.0 / g.xke;
    g.j3oj2 = g.j3 / g.j2;
    return true;
}

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct perturb_gpu_batch {
    int device;
    size_t n, n_pad;
    pb::GravAll grav;
    bool opsmode_a;

    double *d_rec;     // [RF_COUNT][n]      caller order
    double *d_pf;      // [PF_COUNT][n_pad]  class-sorted
    int32_t *d_perm;   // [n_pad]
    int tile_begin[PB_CLS_COUNT];
    int tile_count[PB_CLS_COUNT];
    std::vector<int32_t> init_error;
    std::vector<uint8_t> cls;

    // staging owned by the handle
    cudaStream_t stream;       // construction-time work (kernel 1, sort, pack)
    cudaStream_t last_stream;  // stream of the most recent propagate call
    cudaStream_t copy_stream;  // D2H of host-resident outputs
    double *d_times;
    size_t times_cap;
    double *d_stage[2];
    int32_t *d_stage_err[2];
    size_t stage_cap[2];  // evaluations each staging buffer holds
    cudaEvent_t ev_done[2], ev_free[2];
    cudaEvent_t ev_call;        // end of the most recent call: the next one is ordered after it
    void *h_bounce[2];          // pinned bounce buffers for pageable host outputs
    size_t bounce_cap[2];       // bytes
    pb::HostWorkers *workers;   // scatter threads (created on first use)
    double2 *d_res_nodes;  // resonance node table [resonant slots][res_cap] of 32-byte records
    int *d_res_k0, *d_res_cnt;
    double *d_res_range;
    int res_cap;
    unsigned long long peers_enabled;  // bit d: this device already has peer access to device d
    cudaStream_t aux[PB_CLS_COUNT];  // forked streams, one per orbit class
    cudaEvent_t ev_fork, ev_join[PB_CLS_COUNT];
    int last_launches;
    int cta_sats;  // tuning knob (PERTURB_GPU_CTA_SATS), 0 = automatic
    void *d_summary;  // scratch of perturb_gpu_propagate_summary
    size_t summary_cap;

    perturb_gpu_batch()
        : device(0), n(0), n_pad(0), opsmode_a(false), d_rec(nullptr), d_pf(nullptr),
          d_perm(nullptr), stream(nullptr), last_stream(nullptr), copy_stream(nullptr),
          d_times(nullptr),
          times_cap(0), last_launches(0), cta_sats(0), d_summary(nullptr), summary_cap(0) {
        if (const char *e = std::getenv("PERTURB_GPU_CTA_SATS")) cta_sats = std::atoi(e);
        for (int c = 0; c < PB_CLS_COUNT; ++c) {
            tile_begin[c] = tile_count[c] = 0;
            aux[c] = nullptr;
            ev_join[c] = nullptr;
        }
        ev_fork = nullptr;
        d_res_nodes = nullptr;
        d_res_k0 = d_res_cnt = nullptr;
        d_res_range = nullptr;
        res_cap = 0;
        peers_enabled = 0;
        for (int b = 0; b < 2; ++b) {
            d_stage[b] = nullptr;
            d_stage_err[b] = nullptr;
            stage_cap[b] = 0;
            ev_done[b] = ev_free[b] = nullptr;
            h_bounce[b] = nullptr;
            bounce_cap[b] = 0;
        }
        ev_call = nullptr;
        workers = nullptr;
    }
};

namespace {

void destroy(perturb_gpu_batch *b) {
    if (b == nullptr) return;
    DeviceGuard guard(b->device);
    cudaDeviceSynchronize();
    cudaFree(b->d_rec);
    cudaFree(b->d_pf);
    cudaFree(b->d_perm);
    cudaFree(b->d_times);
    for (int i = 0; i < 2; ++i) {
        cudaFree(b->d_stage[i]);
        cudaFree(b->d_stage_err[i]);
        if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]);
        if (b->ev_free[i]) cudaEventDestroy(b->ev_free[i]);
        if (b->h_bounce[i]) cudaFreeHost(b->h_bounce[i]);
    }
    if (b->ev_call) cudaEventDestroy(b->ev_call);
    delete b->workers;
    for (int c = 0; c < PB_CLS_COUNT; ++c) {
        if (b->aux[c]) cudaStreamDestroy(b->aux[c]);
        if (b->ev_join[c]) cudaEventDestroy(b->ev_join[c]);
    }
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    cudaFree(b->d_summary);
    cudaFree(b->d_res_nodes);
    cudaFree(b->d_res_k0);
    cudaFree(b->d_res_cnt);
    cudaFree(b->d_res_range);
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    delete b;
}

// Classify + sort: stable counting sort by orbit class, every class padded to whole tiles.
int build_sorted_layout(perturb_gpu_batch *b, double *d_pf_stage) {
    const size_t n = b->n;
    size_t count[PB_CLS_COUNT] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) count[std::min<int>(b->cls[i], PB_CLS_COUNT - 1)]++;
    size_t slot = 0;
    size_t first[PB_CLS_COUNT];
    for (int c = 0; c < PB_CLS_COUNT; ++c) {
        first[c] = slot;
        b->tile_begin[c] = static_cast<int>(slot / PB_TILE_SATS);
        const size_t tiles = (count[c] + PB_TILE_SATS - 1) / PB_TILE_SATS;
        b->tile_count[c] = static_cast<int>(tiles);
        slot += tiles * PB_TILE_SATS;
    }
    b->n_pad = slot;
    std::vector<int32_t> perm(b->n_pad, -1);
    size_t cursor[PB_CLS_COUNT];
    for (int c = 0; c < PB_CLS_COUNT; ++c) cursor[c] = first[c];
    for (size_t i = 0; i < n; ++i) {
        const int c = std::min<int>(b->cls[i], PB_CLS_COUNT - 1);
        perm[cursor[c]++] = static_cast<int32_t>(i);
    }
    if (b->n_pad == 0) return PERTURB_GPU_OK;
    PB_CUDA(cudaMalloc(&b->d_perm, b->n_pad * sizeof(int32_t)));
    PB_CUDA(cudaMalloc(&b->d_pf, b->n_pad * PF_COUNT * sizeof(double)));
    PB_CUDA(cudaMemcpyAsync(b->d_perm, perm.data(), b->n_pad * sizeof(int32_t),
                            cudaMemcpyHostToDevice, b->stream));
    pb::launch_pack(d_pf_stage, static_cast<int>(n), b->d_perm, static_cast<int>(b->n_pad),
                    b->d_pf, b->stream);
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(b->stream));
    return PERTURB_GPU_OK;
}

int create_common(size_t n, int grav_model, char opsmode, int device, perturb_gpu_batch **out,
                  perturb_gpu_batch **made) {
    if (out == nullptr) return fail(PERTURB_GPU_ERR_ARG, "out handle pointer is null");
    *out = nullptr;
    if (opsmode != 'i' && opsmode != 'a') {
        return fail(PERTURB_GPU_ERR_ARG, "opsmode must be 'i' or 'a'");
    }
    if (n > static_cast<size_t>(1) << 30) return fail(PERTURB_GPU_ERR_ARG, "batch too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PERTURB_GPU_ERR_NODEV, "no CUDA device available (there is no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(PERTURB_GPU_ERR_ARG, "bad device ordinal");
    perturb_gpu_batch *b = new (std::nothrow) perturb_gpu_batch();
    if (b == nullptr) return fail(PERTURB_GPU_ERR_ALLOC, "host allocation failed");
    b->device = device;
    b->n = n;
    b->opsmode_a = (opsmode == 'a');
    if (!gravity_constants(grav_model, b->grav)) {
        delete b;
        return fail(PERTURB_GPU_ERR_ARG, "unknown gravity model");
    }
    *made = b;
    return PERTURB_GPU_OK;
}

int finish_create(perturb_gpu_batch *b, double *d_pf_stage, uint8_t *d_cls, int32_t *d_err,
                  int32_t *init_error) {
    const size_t n = b->n;
    b->cls.assign(n, 0);
    b->init_error.assign(n, 0);
    if (n > 0) {
        PB_CUDA(cudaMemcpyAsync(b->cls.data(), d_cls, n, cudaMemcpyDeviceToHost, b->stream));
        PB_CUDA(cudaMemcpyAsync(b->init_error.data(), d_err, n * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, b->stream));
        PB_CUDA(cudaStreamSynchronize(b->stream));
    }
    const int st = build_sorted_layout(b, d_pf_stage);
    if (st != PERTURB_GPU_OK) return st;
    if (init_error != nullptr && n > 0) {
        std::memcpy(init_error, b->init_error.data(), n * sizeof(int32_t));
    }
    // Resonance node table: 64 nodes (32 days of grid span) per resonant satellite, less if
    // that would exceed 1 GiB; times outside the table integrate in-line (same bits, slower).
    const size_t res_slots = static_cast<size_t>(b->tile_count[PB_CLS_DEEP_R1] +
                                                 b->tile_count[PB_CLS_DEEP_R2]) * PB_TILE_SATS;
    if (res_slots > 0) {
        size_t cap = 64;
        // one 32-byte record per node: the integrator state (xli, xni) and its rates (xndt, xnddt)
        // (at most 2^25 records, so that record indices fit an int with room to spare)
        while (cap > 4 && 2 * cap * res_slots * sizeof(double2) > (static_cast<size_t>(1) << 30)) cap /= 2;
        PB_CUDA(cudaMalloc(&b->d_res_nodes, 2 * cap * res_slots * sizeof(double2)));
        PB_CUDA(cudaMalloc(&b->d_res_k0, res_slots * sizeof(int)));
        PB_CUDA(cudaMalloc(&b->d_res_cnt, res_slots * sizeof(int)));
        PB_CUDA(cudaMalloc(&b->d_res_range, 4 * sizeof(double)));
        b->res_cap = static_cast<int>(cap);
    }
    return PERTURB_GPU_OK;
}

int create_streams(perturb_gpu_batch *b) {
    PB_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    PB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PB_CUDA(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&b->ev_free[i], cudaEventDisableTiming));
    }
    PB_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    PB_CUDA(cudaEventCreateWithFlags(&b->ev_call, cudaEventDisableTiming));
    PB_CUDA(cudaEventRecord(b->ev_call, b->stream));
    for (int c = 0; c < PB_CLS_COUNT; ++c) {
        PB_CUDA(cudaStreamCreateWithFlags(&b->aux[c], cudaStreamNonBlocking));
        PB_CUDA(cudaEventCreateWithFlags(&b->ev_join[c], cudaEventDisableTiming));
    }
    return PERTURB_GPU_OK;
}

struct ScratchFree {
    void *p[4];
    ScratchFree() { p[0] = p[1] = p[2] = p[3] = nullptr; }
    ~ScratchFree() {
        for (int i = 0; i < 4; ++i) cudaFree(p[i]);
    }
};

struct TextSource {
    const char *lines1, *lines2;
    size_t stride;
    int flavor;
    int32_t *parse_status;  // host, optional
};

int init_impl(const void *src, bool from_records, size_t n, int grav_model, char opsmode,
              int device, perturb_gpu_batch **out, int32_t *init_error,
              const TextSource *text = nullptr) {
    if (src == nullptr && text == nullptr && n > 0) {
        return fail(PERTURB_GPU_ERR_ARG, "input array is null");
    }
    perturb_gpu_batch *b = nullptr;
    int st = create_common(n, grav_model, opsmode, device, out, &b);
    if (st != PERTURB_GPU_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) {
        delete b;
        return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    }
    struct Cleanup {
        perturb_gpu_batch *b;
        ~Cleanup() {
            if (b) destroy(b);
        }
    } cleanup = {b};

    st = create_streams(b);
    if (st != PERTURB_GPU_OK) return st;

    ScratchFree scratch;  // [0] input copy, [1] pf staging, [2] cls, [3] err
    ScratchFree text_scratch;  // [0] lines1, [1] lines2, [2] parse status
    double *d_pf_stage = nullptr;
    uint8_t *d_cls = nullptr;
    int32_t *d_err = nullptr;
    if (n > 0) {
        PB_CUDA(cudaMalloc(&b->d_rec, n * RF_COUNT * sizeof(double)));
        PB_CUDA(cudaMalloc(&scratch.p[1], n * PF_COUNT * sizeof(double)));
        PB_CUDA(cudaMalloc(&scratch.p[2], n));
        PB_CUDA(cudaMalloc(&scratch.p[3], n * sizeof(int32_t)));
        d_pf_stage = static_cast<double *>(scratch.p[1]);
        d_cls = static_cast<uint8_t *>(scratch.p[2]);
        d_err = static_cast<int32_t *>(scratch.p[3]);

        if (from_records) {
            const size_t bytes = n * RF_COUNT * sizeof(double);
            PB_CUDA(cudaMemcpyAsync(b->d_rec, src, bytes, cudaMemcpyDefault, b->stream));
            pb::launch_derive_from_records(static_cast<int>(n), b->d_rec, d_pf_stage, d_cls,
                                           d_err, b->stream);
        } else if (text != nullptr) {
            // TLE text -> numbers on the device, then straight into kernel 1
            const size_t bytes = n * text->stride;
            const char *d_l1 = text->lines1, *d_l2 = text->lines2;
            if (!is_device_pointer(text->lines1)) {
                PB_CUDA(cudaMalloc(&text_scratch.p[0], bytes));
                PB_CUDA(cudaMemcpyAsync(text_scratch.p[0], text->lines1, bytes,
                                        cudaMemcpyHostToDevice, b->stream));
                d_l1 = static_cast<const char *>(text_scratch.p[0]);
            }
            if (!is_device_pointer(text->lines2)) {
                PB_CUDA(cudaMalloc(&text_scratch.p[1], bytes));
                PB_CUDA(cudaMemcpyAsync(text_scratch.p[1], text->lines2, bytes,
                                        cudaMemcpyHostToDevice, b->stream));
                d_l2 = static_cast<const char *>(text_scratch.p[1]);
            }
            PB_CUDA(cudaMalloc(&scratch.p[0], n * sizeof(perturb_gpu_tle)));
            PB_CUDA(cudaMalloc(&text_scratch.p[2], n * sizeof(int32_t)));
            int32_t *d_status = static_cast<int32_t *>(text_scratch.p[2]);
            pb::launch_tle_parse(d_l1, d_l2, text->stride, static_cast<int>(n), text->flavor,
                                 scratch.p[0], d_status, b->stream);
            PB_CUDA(cudaGetLastError());
            if (text->parse_status != nullptr) {
                PB_CUDA(cudaMemcpyAsync(text->parse_status, d_status, n * sizeof(int32_t),
                                        cudaMemcpyDeviceToHost, b->stream));
            }
            pb::launch_sgp4init(scratch.p[0], static_cast<int>(n), b->grav, b->opsmode_a,
                                b->d_rec, d_pf_stage, d_cls, d_err, b->stream);
        } else {
            const void *d_tles = src;
            if (!is_device_pointer(src)) {
                const size_t bytes = n * sizeof(perturb_gpu_tle);
                PB_CUDA(cudaMalloc(&scratch.p[0], bytes));
                PB_CUDA(cudaMemcpyAsync(scratch.p[0], src, bytes, cudaMemcpyHostToDevice,
                                        b->stream));
                d_tles = scratch.p[0];
            }
            pb::launch_sgp4init(d_tles, static_cast<int>(n), b->grav, b->opsmode_a, b->d_rec,
                                d_pf_stage, d_cls, d_err, b->stream);
        }
        PB_CUDA(cudaGetLastError());
    }
    st = finish_create(b, d_pf_stage, d_cls, d_err, init_error);
    if (st != PERTURB_GPU_OK) return st;
    cleanup.b = nullptr;
    *out = b;
    return PERTURB_GPU_OK;
}

int ensure_times(perturb_gpu_batch *b, size_t m) {
    if (b->times_cap >= m) return PERTURB_GPU_OK;
    cudaFree(b->d_times);
    b->d_times = nullptr;
    b->times_cap = 0;
    PB_CUDA(cudaMalloc(&b->d_times, 2 * m * sizeof(double)));
    b->times_cap = m;
    return PERTURB_GPU_OK;
}

int ensure_stage(perturb_gpu_batch *b, int which, size_t evals, size_t doubles_per_eval = 6) {
    // capacity is kept in units of 6-double evaluations (the SoA planes); records need 8
    evals = (evals * doubles_per_eval + 5) / 6;
    if (b->stage_cap[which] >= evals) return PERTURB_GPU_OK;
    cudaFree(b->d_stage[which]);
    cudaFree(b->d_stage_err[which]);
    b->d_stage[which] = nullptr;
    b->d_stage_err[which] = nullptr;
    b->stage_cap[which] = 0;
    PB_CUDA(cudaMalloc(&b->d_stage[which], evals * 6 * sizeof(double)));
    PB_CUDA(cudaMalloc(&b->d_stage_err[which], evals * sizeof(int32_t)));
    b->stage_cap[which] = evals;
    return PERTURB_GPU_OK;
}

void fill_launch(const perturb_gpu_batch *b, pb::PropLaunch &l) {
    l.pf = b->d_pf;
    l.n_pad = b->n_pad;
    l.perm = b->d_perm;
    for (int c = 0; c < PB_CLS_COUNT; ++c) {
        l.tile_begin[c] = b->tile_begin[c];
        l.tile_count[c] = b->tile_count[c];
    }
    l.radiusearthkm = b->grav.radiusearthkm;
    l.xke = b->grav.xke;
    l.j2 = b->grav.j2;
    l.j3oj2 = b->grav.j3oj2;
    l.opsmode_a = b->opsmode_a ? 1 : 0;
    l.cta_sats = b->cta_sats;
    l.res_nodes = b->d_res_nodes;
    l.res_k0 = b->d_res_k0;
    l.res_cnt = b->d_res_cnt;
    l.res_range = b->d_res_range;
    l.res_cap = b->res_cap;
    l.stations = nullptr;
    l.n_stations = 0;
    l.gmst = nullptr;
    l.vis_words = nullptr;
    l.vis_best = nullptr;
    l.pass_out = nullptr;
    l.passes = nullptr;
    l.max_passes = 0;
    l.primaries = nullptr;
    l.n_primaries = 0;
    l.threshold_km = 0.0;
    l.prox_partials = nullptr;
    l.approach_out = nullptr;
    l.out = nullptr;
    l.err = nullptr;
    l.states = nullptr;
    l.elements = nullptr;
    l.summary_partials = nullptr;
    l.summary_out = nullptr;
    l.plane_stride = l.row_stride = l.err_row_stride = 0;
    l.fork = b->ev_fork;
    l.fork_recorded = false;
    for (int c = 0; c < PB_CLS_COUNT; ++c) {
        l.aux[c] = b->aux[c];
        l.join[c] = b->ev_join[c];
    }
}

// ensure_peer_access once per (batch, owner device)
int peer_ready(perturb_gpu_batch *b, int owner) {
    if (owner < 0 || owner == b->device) return PERTURB_GPU_OK;
    if (owner < 64 && ((b->peers_enabled >> owner) & 1ULL)) return PERTURB_GPU_OK;
    const int st = ensure_peer_access(b->device, owner);
    if (st == PERTURB_GPU_OK && owner < 64) b->peers_enabled |= 1ULL << owner;
    return st;
}

int ensure_workers(perturb_gpu_batch *b) {
    if (b->workers != nullptr) return PERTURB_GPU_OK;
    int devices = 1;
    if (cudaGetDeviceCount(&devices) != cudaSuccess) devices = 1;
    unsigned threads = pb::default_host_workers(devices);
    if (const char *e = std::getenv("PERTURB_GPU_HOST_THREADS")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 256) threads = static_cast<unsigned>(v);
    }
    b->workers = new (std::nothrow) pb::HostWorkers(threads);
    if (b->workers == nullptr) return fail(PERTURB_GPU_ERR_ALLOC, "host allocation failed");
    return PERTURB_GPU_OK;
}

// The drop-in call into a REGISTERED (page-locked) array of StateVector records: the copy engine
// writes the caller's memory directly, no bounce copy. The reference's untouched-on-error
// semantics need the failed evaluations' records back as they were: per time-chunk the error
// codes come first (small, through the handle's bounce buffer), the host saves the caller's
// records where a code is 1-4, only then is the chunk's 2-D copy of states enqueued; saved
// position / velocity are put back when everything has landed. `err` may be any host memory
// or NULL.
int propagate_into_registered(perturb_gpu_batch *b, pb::PropLaunch &l, const double *d_ta,
                              const double *d_tb, bool use_jd, size_t m, double *states,
                              int32_t *err, size_t row_stride, cudaStream_t stream) {
    const size_t n = b->n;
    size_t free_b = 0, total_b = 0;
    PB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = std::min<size_t>(free_b / 4, static_cast<size_t>(1) << 30);  // per staging buffer
    for (int i = 0; i < 2; ++i) budget = std::max(budget, b->stage_cap[i] * 52);
    size_t mc = budget / (n * 68);
    if (mc >= 256) mc &= ~static_cast<size_t>(127);
    mc &= ~static_cast<size_t>(1);
    if (mc < 2) mc = 2;
    mc = std::min(mc, (m + 1) & ~static_cast<size_t>(1));
    const size_t nchunks = (m + mc - 1) / mc;
    for (int i = 0; i < (nchunks > 1 ? 2 : 1); ++i) {
        int st = ensure_stage(b, i, n * mc, 9);  // records + codes in one block
        if (st != PERTURB_GPU_OK) return st;
        const size_t need = n * mc * sizeof(int32_t);
        if (b->bounce_cap[i] < need) {
            if (b->h_bounce[i]) cudaFreeHost(b->h_bounce[i]);
            b->h_bounce[i] = nullptr;
            b->bounce_cap[i] = 0;
            if (cudaHostAlloc(&b->h_bounce[i], need, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return fail(PERTURB_GPU_ERR_ALLOC, "pinned bounce buffer allocation failed");
            }
            b->bounce_cap[i] = need;
        }
    }
    struct Saved {
        size_t at;  // record index i * row_stride + k
        double pv[6];
    };
    std::vector<Saved> saved;
    std::mutex saved_mu;
    int st = ensure_workers(b);
    if (st != PERTURB_GPU_OK) return st;
    auto launch_chunk = [&](size_t c) -> int {
        const int buf = static_cast<int>(c % 2);
        const size_t k0 = c * mc, mk = std::min(mc, m - k0), ld = (mk + 1) & ~static_cast<size_t>(1);
        if (c >= 2) PB_CUDA(cudaStreamWaitEvent(stream, b->ev_free[buf], 0));
        l.ta = d_ta + k0;
        l.tb = use_jd ? d_tb + k0 : nullptr;
        l.m = static_cast<int>(mk);
        l.out = nullptr;
        l.states = b->d_stage[buf];
        l.row_stride = ld;
        l.plane_stride = n * ld;
        l.err = reinterpret_cast<int32_t *>(b->d_stage[buf] + n * ld * 8);
        l.err_row_stride = ld;
        const int launched = pb::launch_sgp4_propagate(l, stream);
        if (launched < 0) {
            return fail(PERTURB_GPU_ERR_ARG, "satellites x times too large for one call: chunk the time grid");
        }
        b->last_launches += launched;
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaEventRecord(b->ev_done[buf], stream));
        PB_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_done[buf], 0));
        PB_CUDA(cudaMemcpyAsync(b->h_bounce[buf], l.err, n * ld * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, b->copy_stream));
        PB_CUDA(cudaEventRecord(b->ev_done[buf], b->copy_stream));  // now: "codes are on the host"
        return PERTURB_GPU_OK;
    };
    st = launch_chunk(0);
    if (st != PERTURB_GPU_OK) return st;
    for (size_t c = 0; c < nchunks; ++c) {
        const int buf = static_cast<int>(c % 2);
        const size_t k0 = c * mc, mk = std::min(mc, m - k0), ld = (mk + 1) & ~static_cast<size_t>(1);
        if (c + 1 < nchunks) {  // the next chunk computes while this one's codes are examined
            st = launch_chunk(c + 1);
            if (st != PERTURB_GPU_OK) return st;
        }
        PB_CUDA(cudaEventSynchronize(b->ev_done[buf]));
        const int32_t *codes = static_cast<const int32_t *>(b->h_bounce[buf]);
        b->workers->parallel_for(n, [&](size_t lo, size_t hi) {
            std::vector<Saved> mine;
            for (size_t i = lo; i < hi; ++i) {
                const int32_t *ei = codes + i * ld;
                if (err != nullptr) std::memcpy(err + i * row_stride + k0, ei, mk * sizeof(int32_t));
                for (size_t k = 0; k < mk; ++k) {
                    if (ei[k] != 0 && ei[k] != 6) {
                        Saved sv;
                        sv.at = i * row_stride + k0 + k;
                        std::memcpy(sv.pv, states + sv.at * 8 + 2, sizeof(sv.pv));
                        mine.push_back(sv);
                    }
                }
            }
            if (!mine.empty()) {
                std::lock_guard<std::mutex> lock(saved_mu);
                saved.insert(saved.end(), mine.begin(), mine.end());
            }
        });
        PB_CUDA(cudaMemcpy2DAsync(states + k0 * 8, row_stride * 64, b->d_stage[buf], ld * 64, mk * 64, n,
                                  cudaMemcpyDeviceToHost, b->copy_stream));
        PB_CUDA(cudaEventRecord(b->ev_free[buf], b->copy_stream));
    }
    PB_CUDA(cudaStreamSynchronize(b->copy_stream));
    for (const Saved &sv : saved) std::memcpy(states + sv.at * 8 + 2, sv.pv, sizeof(sv.pv));
    PB_CUDA(cudaEventRecord(b->ev_call, stream));
    return PERTURB_GPU_OK;
}

int propagate_impl(perturb_gpu_batch *b, const double *ta, const double *tb, bool use_jd,
                   size_t m, double *posvel, int32_t *err, size_t row_stride,
                   size_t plane_stride, void *stream_arg, double *elements = nullptr,
                   bool records = false, bool keep_failed = false) {
    // records: `posvel` is an array of perturb_gpu_state (8 doubles each) indexed
    // [sat * row_stride + k]; plane_stride is unused.
    // keep_failed (records in host memory only): position / velocity of evaluations with codes
    // 1-4 are left as the caller had them, like the reference (sgp4.cpp:1865,1878,1943,1995).
    if (b == nullptr) return fail(PERTURB_GPU_ERR_ARG, "batch is null");
    b->last_launches = 0;
    if (m == 0 || b->n == 0) return PERTURB_GPU_OK;
    if (ta == nullptr || (use_jd && tb == nullptr) || posvel == nullptr ||
        (err == nullptr && !keep_failed)) {
        return fail(PERTURB_GPU_ERR_ARG, "null time or output pointer");
    }
    if (row_stride < m || (!records && plane_stride < b->n * row_stride)) {
        return fail(PERTURB_GPU_ERR_ARG, "row_stride < m or plane_stride < n * row_stride");
    }
    if (records && is_device_pointer(posvel) && (reinterpret_cast<uintptr_t>(posvel) & 15u) != 0) {
        return fail(PERTURB_GPU_ERR_ARG, "device-resident states must be 16-byte aligned");
    }
    if (m > static_cast<size_t>(1) << 30) return fail(PERTURB_GPU_ERR_ARG, "time grid too long");
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    // NULL is the (legacy) default stream, as in the CUDA runtime API.
    cudaStream_t stream = static_cast<cudaStream_t>(stream_arg);
    b->last_stream = stream;
    // One call at a time per handle: the scratch below (time grid, node table, staging) is the
    // handle's, so a call on another stream is ordered after the previous one.
    PB_CUDA(cudaStreamWaitEvent(stream, b->ev_call, 0));

    // ---- time grid on the device ----
    // (a grid that lives on another device is copied like a host one: it is read n / 32 times)
    const bool ta_dev = is_local_device_pointer(ta, b->device);
    const bool tb_dev = use_jd ? is_local_device_pointer(tb, b->device) : true;
    const double *d_ta = ta, *d_tb = tb;
    if (!ta_dev || !tb_dev) {
        int st = ensure_times(b, m);
        if (st != PERTURB_GPU_OK) return st;
        if (!ta_dev) {
            PB_CUDA(cudaMemcpyAsync(b->d_times, ta, m * sizeof(double), cudaMemcpyDefault,
                                    stream));
            d_ta = b->d_times;
        }
        if (use_jd && !tb_dev) {
            PB_CUDA(cudaMemcpyAsync(b->d_times + m, tb, m * sizeof(double), cudaMemcpyDefault,
                                    stream));
            d_tb = b->d_times + m;
        }
    }

    pb::PropLaunch l;
    fill_launch(b, l);
    l.use_jd = use_jd ? 1 : 0;
    l.elements = elements;
    l.states = nullptr;
    l.summary_partials = nullptr;
    l.summary_out = nullptr;
    l.n = b->n;
    if (elements != nullptr && !(is_device_pointer(elements) && is_device_pointer(posvel) &&
                                 is_device_pointer(err))) {
        return fail(PERTURB_GPU_ERR_ARG, "propagate_elements needs device output buffers");
    }

    const bool out_dev = is_device_pointer(posvel);
    const bool err_dev = (err != nullptr) ? is_device_pointer(err) : out_dev;
    if (out_dev != err_dev) {
        return fail(PERTURB_GPU_ERR_ARG, "posvel and err must both be host or both be device");
    }
    if (keep_failed && (out_dev || !records)) {
        return fail(PERTURB_GPU_ERR_ARG, "keep_failed needs StateVector records in host memory");
    }

    if (out_dev) {
        // device-resident gather: the buffers may belong to another GPU of this process
        for (const void *q : {static_cast<const void *>(posvel), static_cast<const void *>(err),
                              static_cast<const void *>(elements)}) {
            const int st = peer_ready(b, pointer_device(q));
            if (st != PERTURB_GPU_OK) return st;
        }
        l.ta = d_ta;
        l.tb = d_tb;
        l.m = static_cast<int>(m);
        l.out = records ? nullptr : posvel;
        l.states = records ? posvel : nullptr;
        l.plane_stride = plane_stride;
        l.row_stride = row_stride;
        l.err = err;
        l.err_row_stride = row_stride;
        b->last_launches = pb::launch_sgp4_propagate(l, stream);
        if (b->last_launches < 0) {
            b->last_launches = 0;
            return fail(PERTURB_GPU_ERR_ARG, "satellites x times too large for one call: chunk the time grid");
        }
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaEventRecord(b->ev_call, stream));
        return PERTURB_GPU_OK;
    }

    // ---- host outputs: time-chunked, double-buffered device staging + async D2H ----
    // Pinned destinations are written by the copy engine directly (asynchronous call);
    // pageable ones (std::vector, numpy) go through pinned bounce buffers that host threads
    // scatter into the caller's arrays while the next chunk is computed and copied (the call
    // returns when the results are in place).
    const bool direct = !keep_failed && is_pinned_host_pointer(posvel) && is_pinned_host_pointer(err);
    if (keep_failed && is_pinned_host_pointer(posvel)) {
        // (records in registered memory: straight DMA, failed evaluations saved and restored)
        try {
            return propagate_into_registered(b, l, d_ta, d_tb, use_jd, m, posvel, err, row_stride, stream);
        } catch (const std::bad_alloc &) {
            return fail(PERTURB_GPU_ERR_ALLOC, "host allocation failed");
        }
    }
    size_t free_b = 0, total_b = 0;
    PB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t dpe = records ? 8 : 6;  // doubles per evaluation in the staging buffer
    const size_t bytes_per_eval = dpe * sizeof(double) + sizeof(int32_t);
    size_t budget = std::min<size_t>(free_b / 4, static_cast<size_t>(8) << 30);  // per buffer
    for (int i = 0; i < 2; ++i) budget = std::max(budget, b->stage_cap[i] * bytes_per_eval);
    // Pinned destinations: ONE chunk by default, i.e. the kernel (2 % of the step: PCIe 55 GB/s
    // against 2.5 TB/s of output) followed by one linear copy per buffer. Time-chunks would hide
    // the kernel behind the previous chunk's copy, but their copies are 2-D (row segments of the
    // caller's [n][m] planes) and those run at 39-50 GB/s on these boxes where the linear copy
    // reaches 54 of the raw 57 (profiles/r2n_e2e_chunks_*.jsonl; PERTURB_GPU_HOST_CHUNKS overrides,
    // and grids whose staging does not fit the budget are chunked anyway).
    const size_t total_bytes = bytes_per_eval * b->n * m;
    if (direct && total_bytes > (static_cast<size_t>(256) << 20)) {
        size_t want = 1;
        if (const char *e = std::getenv("PERTURB_GPU_HOST_CHUNKS")) {
            const int v = std::atoi(e);
            if (v >= 1 && v <= 64) want = static_cast<size_t>(v);
        }
        if (want > 1) {
            budget = std::min(budget, std::max<size_t>((total_bytes + want - 1) / want + bytes_per_eval * b->n * 2,
                                                       static_cast<size_t>(16) << 20));
        }
    }
    if (!direct) budget = std::min(budget, static_cast<size_t>(128) << 20);
    const size_t m_even = (m + 1) & ~static_cast<size_t>(1);
    size_t fit = budget / (bytes_per_eval * b->n);  // times one staging buffer can hold
    if (!direct && fit >= 256) fit &= ~static_cast<size_t>(127);
    fit &= ~static_cast<size_t>(1);
    if (fit < 2) fit = 2;
    // Balanced chunks (not "full, full, ..., sliver"): short rows make poor 2-D copies.
    const size_t nchunks = (m_even + fit - 1) / fit;
    size_t mc = ((m + nchunks - 1) / nchunks + 1) & ~static_cast<size_t>(1);  // even
    if (mc >= 256 && nchunks > 1) mc = std::min(fit, (mc + 127) & ~static_cast<size_t>(127));
    const size_t mc_pad = mc;
    const int nbuf = nchunks > 1 ? 2 : 1;
    for (int i = 0; i < nbuf; ++i) {
        // (bounce chunks keep their error codes behind the states in the same block)
        int st = ensure_stage(b, i, b->n * mc_pad, direct ? dpe : dpe + 1);
        if (st != PERTURB_GPU_OK) return st;
        if (!direct) {
            const size_t need = b->n * mc_pad * bytes_per_eval;
            if (b->bounce_cap[i] < need) {
                if (b->h_bounce[i]) cudaFreeHost(b->h_bounce[i]);
                b->h_bounce[i] = nullptr;
                b->bounce_cap[i] = 0;
                if (cudaHostAlloc(&b->h_bounce[i], need, cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError();
                    return fail(PERTURB_GPU_ERR_ALLOC, "pinned bounce buffer allocation failed");
                }
                b->bounce_cap[i] = need;
            }
        }
    }
    if (!direct) {
        const int st = ensure_workers(b);
        if (st != PERTURB_GPU_OK) return st;
    }

    // Bounce chunk -> the caller's arrays (rows of one chunk, host threads).
    struct Pending {
        bool valid;
        int buf;
        size_t k0, mk, ld;
    } pending = {false, 0, 0, 0, 0};
    auto scatter = [&](const Pending &pc) {
        const size_t n = b->n;
        const char *src = static_cast<const char *>(b->h_bounce[pc.buf]);
        const int32_t *src_err = reinterpret_cast<const int32_t *>(src + n * pc.ld * dpe * sizeof(double));
        const size_t k0 = pc.k0, mk = pc.mk, ld = pc.ld;
        b->workers->parallel_for(n, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const int32_t *ei = src_err + i * ld;
                if (err != nullptr) std::memcpy(err + i * row_stride + k0, ei, mk * sizeof(int32_t));
                if (records) {
                    const char *si = src + i * ld * 64;
                    char *di = reinterpret_cast<char *>(posvel) + (i * row_stride + k0) * 64;
                    bool clean = true;
                    if (keep_failed) {
                        for (size_t k = 0; k < mk; ++k) clean = clean && (ei[k] == 0 || ei[k] == 6);
                    }
                    if (clean) {
                        copy_streaming(di, si, mk * 64);
                    } else {
                        for (size_t k = 0; k < mk; ++k) {  // epoch always, r / v for codes 0 and 6
                            const bool written = (ei[k] == 0 || ei[k] == 6);
                            std::memcpy(di + k * 64, si + k * 64, written ? 64 : 16);
                        }
                    }
                } else {
                    for (size_t pl = 0; pl < 6; ++pl) {
                        copy_streaming(posvel + pl * plane_stride + i * row_stride + k0,
                                       src + (pl * n * ld + i * ld) * sizeof(double), mk * sizeof(double));
                    }
                }
            }
            streaming_fence();
        });
    };

    for (size_t c = 0; c < nchunks; ++c) {
        const int buf = static_cast<int>(c % 2);
        const size_t k0 = c * mc;
        if (k0 >= m) break;
        const size_t mk = std::min(mc, m - k0);
        if (c >= 2) PB_CUDA(cudaStreamWaitEvent(stream, b->ev_free[buf], 0));
        l.ta = d_ta + k0;
        l.tb = use_jd ? d_tb + k0 : nullptr;
        l.m = static_cast<int>(mk);
        l.out = records ? nullptr : b->d_stage[buf];
        l.states = records ? b->d_stage[buf] : nullptr;
        // one chunk covering whole contiguous caller rows: stage with the caller's own
        // strides so that the D2H is a single linear copy per buffer
        const bool linear = direct && (nchunks == 1) && (row_stride == m) && (m % 2 == 0);
        // bounce chunks are packed (row length = the chunk's own, rounded up to even)
        l.row_stride = linear ? m : direct ? mc_pad : ((mk + 1) & ~static_cast<size_t>(1));
        l.plane_stride = b->n * l.row_stride;
        l.err = direct ? b->d_stage_err[buf]
                       : reinterpret_cast<int32_t *>(b->d_stage[buf] + b->n * l.row_stride * dpe);
        l.err_row_stride = l.row_stride;
        const int launched = pb::launch_sgp4_propagate(l, stream);
        if (launched < 0) {
            return fail(PERTURB_GPU_ERR_ARG, "satellites x times too large for one call: chunk the time grid");
        }
        b->last_launches += launched;
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaEventRecord(b->ev_done[buf], stream));
        PB_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_done[buf], 0));
        if (!direct) {
            // states and codes of the chunk are one contiguous block of the staging buffer
            PB_CUDA(cudaMemcpyAsync(b->h_bounce[buf], b->d_stage[buf],
                                    b->n * l.row_stride * bytes_per_eval, cudaMemcpyDeviceToHost,
                                    b->copy_stream));
            PB_CUDA(cudaEventRecord(b->ev_free[buf], b->copy_stream));
            if (pending.valid) {  // the previous chunk, while this one is computed and copied
                PB_CUDA(cudaEventSynchronize(b->ev_free[pending.buf]));
                scatter(pending);
            }
            pending = {true, buf, k0, mk, l.row_stride};
            continue;
        }
        if (records) {
            // rows of 64-byte records: one 2-D copy (or a linear one when the rows are whole)
            const size_t rb = 8 * sizeof(double);
            if (linear) {
                PB_CUDA(cudaMemcpyAsync(posvel, b->d_stage[buf], b->n * m * rb,
                                        cudaMemcpyDeviceToHost, b->copy_stream));
                PB_CUDA(cudaMemcpyAsync(err, b->d_stage_err[buf], b->n * m * sizeof(int32_t),
                                        cudaMemcpyDeviceToHost, b->copy_stream));
            } else {
                PB_CUDA(cudaMemcpy2DAsync(posvel + k0 * 8, row_stride * rb, b->d_stage[buf],
                                          mc_pad * rb, mk * rb, b->n, cudaMemcpyDeviceToHost,
                                          b->copy_stream));
                PB_CUDA(cudaMemcpy2DAsync(err + k0, row_stride * sizeof(int32_t),
                                          b->d_stage_err[buf], mc_pad * sizeof(int32_t),
                                          mk * sizeof(int32_t), b->n, cudaMemcpyDeviceToHost,
                                          b->copy_stream));
            }
        } else if (linear) {
            if (plane_stride == b->n * m) {
                PB_CUDA(cudaMemcpyAsync(posvel, b->d_stage[buf], 6 * b->n * m * sizeof(double),
                                        cudaMemcpyDeviceToHost, b->copy_stream));
            } else {
                for (int pl = 0; pl < 6; ++pl) {
                    PB_CUDA(cudaMemcpyAsync(posvel + pl * plane_stride,
                                            b->d_stage[buf] + pl * l.plane_stride,
                                            b->n * m * sizeof(double), cudaMemcpyDeviceToHost,
                                            b->copy_stream));
                }
            }
            PB_CUDA(cudaMemcpyAsync(err, b->d_stage_err[buf], b->n * m * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, b->copy_stream));
        } else {
            for (int pl = 0; pl < 6; ++pl) {
                PB_CUDA(cudaMemcpy2DAsync(posvel + pl * plane_stride + k0,
                                          row_stride * sizeof(double),
                                          b->d_stage[buf] + pl * l.plane_stride,
                                          mc_pad * sizeof(double), mk * sizeof(double), b->n,
                                          cudaMemcpyDeviceToHost, b->copy_stream));
            }
            PB_CUDA(cudaMemcpy2DAsync(err + k0, row_stride * sizeof(int32_t),
                                      b->d_stage_err[buf], mc_pad * sizeof(int32_t),
                                      mk * sizeof(int32_t), b->n, cudaMemcpyDeviceToHost,
                                      b->copy_stream));
        }
        PB_CUDA(cudaEventRecord(b->ev_free[buf], b->copy_stream));
    }
    if (pending.valid) {
        PB_CUDA(cudaEventSynchronize(b->ev_free[pending.buf]));
        scatter(pending);
    }
    // The caller's stream must see the copies as part of this call.
    const int last = static_cast<int>((nchunks - 1) % 2);
    PB_CUDA(cudaStreamWaitEvent(stream, b->ev_free[last], 0));
    if (nchunks > 1) PB_CUDA(cudaStreamWaitEvent(stream, b->ev_free[1 - last], 0));
    PB_CUDA(cudaEventRecord(b->ev_call, stream));
    return PERTURB_GPU_OK;
}

}  // namespace

extern "C" {

int perturb_gpu_init(const perturb_gpu_tle *tles, size_t n, int grav_model, char opsmode,
                     int device, perturb_gpu_batch **out, int32_t *init_error) {
    return init_impl(tles, false, n, grav_model, opsmode, device, out, init_error);
}

int perturb_gpu_init_from_text(const char *lines1, const char *lines2, size_t line_stride,
                               size_t n, int flavor, int grav_model, char opsmode, int device,
                               perturb_gpu_batch **out, int32_t *init_error,
                               int32_t *parse_status) {
    if ((n > 0 && (lines1 == nullptr || lines2 == nullptr)) || line_stride < 69 ||
        (flavor != 0 && flavor != 1)) {
        return fail(PERTURB_GPU_ERR_ARG, "bad TLE text arguments");
    }
    const TextSource text = {lines1, lines2, line_stride, flavor, parse_status};
    return init_impl(nullptr, false, n, grav_model, opsmode, device, out, init_error, &text);
}

int perturb_gpu_init_records(const double *records, size_t n, int grav_model, char opsmode,
                             int device, perturb_gpu_batch **out, int32_t *init_error) {
    return init_impl(records, true, n, grav_model, opsmode, device, out, init_error);
}

void perturb_gpu_free(perturb_gpu_batch *batch) { destroy(batch); }

int perturb_gpu_propagate(perturb_gpu_batch *batch, const double *jd, const double *jd_frac,
                          size_t m, double *posvel, int32_t *err, size_t row_stride,
                          size_t plane_stride, void *stream) {
    return propagate_impl(batch, jd, jd_frac, true, m, posvel, err, row_stride, plane_stride,
                          stream);
}

int perturb_gpu_propagate_from_epoch(perturb_gpu_batch *batch, const double *mins, size_t m,
                                     double *posvel, int32_t *err, size_t row_stride,
                                     size_t plane_stride, void *stream) {
    return propagate_impl(batch, mins, nullptr, false, m, posvel, err, row_stride, plane_stride,
                          stream);
}

int perturb_gpu_propagate_states(perturb_gpu_batch *batch, const double *jd,
                                 const double *jd_frac, size_t m, perturb_gpu_state *states,
                                 int32_t *err, size_t row_stride, void *stream) {
    static_assert(sizeof(perturb_gpu_state) == 8 * sizeof(double), "StateVector is 64 bytes");
    return propagate_impl(batch, jd, jd_frac, jd_frac != nullptr, m,
                          reinterpret_cast<double *>(states), err, row_stride, 0, stream, nullptr,
                          true);
}

int perturb_gpu_propagate_into(perturb_gpu_batch *batch, const double *jd, const double *jd_frac,
                               size_t m, perturb_gpu_state *states, int32_t *err,
                               size_t row_stride) {
    const int st = propagate_impl(batch, jd, jd_frac, jd_frac != nullptr, m,
                                  reinterpret_cast<double *>(states), err, row_stride, 0, nullptr,
                                  nullptr, true, true);
    if (st != PERTURB_GPU_OK) return st;
    return perturb_gpu_synchronize(batch);
}

int perturb_gpu_propagate_elements(perturb_gpu_batch *batch, const double *jd,
                                   const double *jd_frac, size_t m, double *posvel,
                                   int32_t *err, double *elements, size_t row_stride,
                                   size_t plane_stride, void *stream) {
    if (elements == nullptr) return fail(PERTURB_GPU_ERR_ARG, "elements is null");
    if (batch == nullptr) return fail(PERTURB_GPU_ERR_ARG, "batch is null");
    const bool all_dev = is_device_pointer(elements) && is_device_pointer(posvel) && is_device_pointer(err);
    if (all_dev || m == 0 || batch->n == 0) {
        return propagate_impl(batch, jd, jd_frac, jd_frac != nullptr, m, posvel, err, row_stride,
                              plane_stride, stream, elements);
    }
    if (is_device_pointer(elements) || is_device_pointer(posvel) || is_device_pointer(err)) {
        return fail(PERTURB_GPU_ERR_ARG, "posvel, err and elements must all be host or all be device memory");
    }
    if (posvel == nullptr || err == nullptr || jd == nullptr || row_stride < m ||
        plane_stride < batch->n * row_stride) {
        return fail(PERTURB_GPU_ERR_ARG, "null pointer, row_stride < m or plane_stride < n * row_stride");
    }
    // Host outputs: time-chunks through device buffers of this call (13 planes + codes, at most
    // ~256 MiB), copied out plane by plane; returns when the results are in place.
    DeviceGuard guard(batch->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    const size_t n = batch->n;
    const size_t per_time = n * (13 * sizeof(double) + sizeof(int32_t));
    size_t mc = std::max<size_t>(2, (static_cast<size_t>(256) << 20) / per_time) & ~static_cast<size_t>(1);
    mc = std::min(mc, (m + 1) & ~static_cast<size_t>(1));
    ScratchFree scratch;  // [0] posvel + elements planes, [1] codes, [2] time grid
    PB_CUDA(cudaMalloc(&scratch.p[0], 13 * n * mc * sizeof(double)));
    PB_CUDA(cudaMalloc(&scratch.p[1], n * mc * sizeof(int32_t)));
    PB_CUDA(cudaMalloc(&scratch.p[2], 2 * m * sizeof(double)));
    double *d_pv = static_cast<double *>(scratch.p[0]);
    double *d_el = d_pv + 6 * n * mc;
    int32_t *d_err = static_cast<int32_t *>(scratch.p[1]);
    double *d_ta = static_cast<double *>(scratch.p[2]), *d_tb = d_ta + m;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PB_CUDA(cudaMemcpyAsync(d_ta, jd, m * sizeof(double), cudaMemcpyDefault, s));
    if (jd_frac != nullptr) {
        PB_CUDA(cudaMemcpyAsync(d_tb, jd_frac, m * sizeof(double), cudaMemcpyDefault, s));
    }
    int launches = 0;
    for (size_t k0 = 0; k0 < m; k0 += mc) {
        const size_t mk = std::min(mc, m - k0);
        const int st = propagate_impl(batch, d_ta + k0, jd_frac != nullptr ? d_tb + k0 : nullptr,
                                      jd_frac != nullptr, mk, d_pv, d_err, mc, n * mc, stream, d_el);
        if (st != PERTURB_GPU_OK) return st;
        launches += batch->last_launches;
        for (int pl = 0; pl < 13; ++pl) {
            double *dst = (pl < 6 ? posvel + pl * plane_stride : elements + (pl - 6) * plane_stride) + k0;
            PB_CUDA(cudaMemcpy2DAsync(dst, row_stride * sizeof(double), d_pv + pl * n * mc,
                                      mc * sizeof(double), mk * sizeof(double), n,
                                      cudaMemcpyDeviceToHost, s));
        }
        PB_CUDA(cudaMemcpy2DAsync(err + k0, row_stride * sizeof(int32_t), d_err, mc * sizeof(int32_t),
                                  mk * sizeof(int32_t), n, cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaStreamSynchronize(s));
    }
    batch->last_launches = launches;
    return PERTURB_GPU_OK;
}

int perturb_gpu_propagate_summary(perturb_gpu_batch *b, const double *jd, const double *jd_frac,
                                  size_t m, perturb_gpu_summary *out, void *stream_arg) {
    if (b == nullptr) return fail(PERTURB_GPU_ERR_ARG, "batch is null");
    b->last_launches = 0;
    if (b->n == 0) return PERTURB_GPU_OK;
    if (jd == nullptr || out == nullptr || m == 0 || m > (static_cast<size_t>(1) << 30)) {
        return fail(PERTURB_GPU_ERR_ARG, "null pointer or empty / oversized time grid");
    }
    const bool use_jd = jd_frac != nullptr;
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_arg);
    b->last_stream = stream;

    const double *d_ta = jd, *d_tb = jd_frac;
    const bool ta_dev = is_local_device_pointer(jd, b->device);
    const bool tb_dev = use_jd ? is_local_device_pointer(jd_frac, b->device) : true;
    if (!ta_dev || !tb_dev) {
        int st = ensure_times(b, m);
        if (st != PERTURB_GPU_OK) return st;
        if (!ta_dev) {
            PB_CUDA(cudaMemcpyAsync(b->d_times, jd, m * sizeof(double), cudaMemcpyDefault,
                                    stream));
            d_ta = b->d_times;
        }
        if (use_jd && !tb_dev) {
            PB_CUDA(cudaMemcpyAsync(b->d_times + m, jd_frac, m * sizeof(double),
                                    cudaMemcpyDefault, stream));
            d_tb = b->d_times + m;
        }
    }
    // scratch: partial records [n][chunks] + (for host output) the folded records [n]
    const size_t chunks = pb::summary_chunks(m);
    const size_t need = (b->n * chunks + b->n) * sizeof(perturb_gpu_summary);
    if (b->summary_cap < need) {
        cudaFree(b->d_summary);
        b->d_summary = nullptr;
        b->summary_cap = 0;
        PB_CUDA(cudaMalloc(&b->d_summary, need));
        b->summary_cap = need;
    }
    perturb_gpu_summary *partials = static_cast<perturb_gpu_summary *>(b->d_summary);
    perturb_gpu_summary *folded = partials + b->n * chunks;
    const bool out_dev = is_device_pointer(out);
    if (out_dev) {
        const int st = peer_ready(b, pointer_device(out));
        if (st != PERTURB_GPU_OK) return st;
    }

    pb::PropLaunch l;
    fill_launch(b, l);
    l.use_jd = use_jd ? 1 : 0;
    l.ta = d_ta;
    l.tb = d_tb;
    l.m = static_cast<int>(m);
    l.out = nullptr;
    l.err = nullptr;
    l.plane_stride = l.row_stride = l.err_row_stride = 0;
    l.elements = nullptr;
    l.summary_partials = partials;
    l.summary_out = out_dev ? out : folded;
    l.n = b->n;
    b->last_launches = pb::launch_sgp4_propagate(l, stream);
    if (b->last_launches < 0) {
        b->last_launches = 0;
        return fail(PERTURB_GPU_ERR_ARG, "satellites x times too large for one call: chunk the time grid");
    }
    PB_CUDA(cudaGetLastError());
    if (!out_dev) {
        PB_CUDA(cudaMemcpyAsync(out, folded, b->n * sizeof(perturb_gpu_summary),
                                cudaMemcpyDeviceToHost, stream));
    }
    return PERTURB_GPU_OK;
}

namespace {

// The handle's general device scratch (shared by the fused consumers; one call at a time).
int ensure_scratch(perturb_gpu_batch *b, size_t bytes) {
    if (b->summary_cap >= bytes) return PERTURB_GPU_OK;
    cudaFree(b->d_summary);
    b->d_summary = nullptr;
    b->summary_cap = 0;
    PB_CUDA(cudaMalloc(&b->d_summary, bytes));
    b->summary_cap = bytes;
    return PERTURB_GPU_OK;
}

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// Time grid of a fused-consumer call on the device (copies host grids into the handle's buffer).
int stage_times(perturb_gpu_batch *b, const double *jd, const double *jd_frac, size_t m,
                cudaStream_t stream, const double **d_ta, const double **d_tb) {
    const bool use_jd = jd_frac != nullptr;
    *d_ta = jd;
    *d_tb = jd_frac;
    const bool ta_dev = is_local_device_pointer(jd, b->device);
    const bool tb_dev = use_jd ? is_local_device_pointer(jd_frac, b->device) : true;
    if (ta_dev && tb_dev) return PERTURB_GPU_OK;
    const int st = ensure_times(b, m);
    if (st != PERTURB_GPU_OK) return st;
    if (!ta_dev) {
        PB_CUDA(cudaMemcpyAsync(b->d_times, jd, m * sizeof(double), cudaMemcpyDefault, stream));
        *d_ta = b->d_times;
    }
    if (use_jd && !tb_dev) {
        PB_CUDA(cudaMemcpyAsync(b->d_times + m, jd_frac, m * sizeof(double), cudaMemcpyDefault,
                                stream));
        *d_tb = b->d_times + m;
    }
    return PERTURB_GPU_OK;
}

}  // namespace

int perturb_gpu_propagate_visibility(perturb_gpu_batch *b, const double *jd,
                                     const double *jd_frac, size_t m,
                                     const perturb_gpu_station *stations, size_t n_stations,
                                     perturb_gpu_pass_summary *out, int32_t *passes,
                                     size_t max_passes, void *stream_arg) {
    static_assert(sizeof(perturb_gpu_pass_summary) == 32, "pass summary is 32 bytes");
    if (b == nullptr) return fail(PERTURB_GPU_ERR_ARG, "batch is null");
    b->last_launches = 0;
    if (b->n == 0 || n_stations == 0) return PERTURB_GPU_OK;
    if (jd == nullptr || jd_frac == nullptr || stations == nullptr || out == nullptr || m == 0 ||
        m > (static_cast<size_t>(1) << 30) || n_stations > 64 || max_passes > 4096 ||
        (passes != nullptr && max_passes == 0)) {
        return fail(PERTURB_GPU_ERR_ARG, "visibility: null pointer, empty / oversized grid, more than "
                                         "64 stations, or passes without max_passes");
    }
    if (is_device_pointer(stations)) return fail(PERTURB_GPU_ERR_ARG, "stations must be host memory");
    const bool out_dev = is_device_pointer(out);
    if (passes != nullptr && is_device_pointer(passes) != out_dev) {
        return fail(PERTURB_GPU_ERR_ARG, "out and passes must both be host or both be device memory");
    }
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_arg);
    b->last_stream = stream;
    PB_CUDA(cudaStreamWaitEvent(stream, b->ev_call, 0));
    const double *d_ta, *d_tb;
    int st = stage_times(b, jd, jd_frac, m, stream, &d_ta, &d_tb);
    if (st != PERTURB_GPU_OK) return st;

    // stations on the ellipsoid of the batch's gravity model -> Earth-fixed position + zenith
    const double flat = (b->grav.radiusearthkm == 6378.137) ? 1.0 / 298.257223563 : 1.0 / 298.26;
    const double e2 = flat * (2.0 - flat), a = b->grav.radiusearthkm;
    std::vector<double> packed(8 * n_stations);
    for (size_t s = 0; s < n_stations; ++s) {
        const double sl = std::sin(stations[s].latitude_rad), cl = std::cos(stations[s].latitude_rad);
        const double so = std::sin(stations[s].longitude_rad), co = std::cos(stations[s].longitude_rad);
        const double nn = a / std::sqrt(1.0 - e2 * sl * sl), h = stations[s].altitude_km;
        double *q = &packed[8 * s];
        q[0] = (nn + h) * cl * co;
        q[1] = (nn + h) * cl * so;
        q[2] = (nn * (1.0 - e2) + h) * sl;
        q[3] = cl * co;
        q[4] = cl * so;
        q[5] = sl;
        q[6] = std::sin(stations[s].min_elevation_rad);
        q[7] = 0.0;
    }
    const size_t chunks = pb::summary_chunks(m), pairs = b->n * n_stations;
    const size_t o_st = 0, o_gm = align256(o_st + packed.size() * sizeof(double));
    const size_t o_words = align256(o_gm + 2 * m * sizeof(double));
    const size_t o_best = align256(o_words + pairs * chunks * sizeof(uint32_t));
    const size_t o_out = align256(o_best + pairs * chunks * sizeof(double2));
    const size_t o_pass = align256(o_out + pairs * sizeof(perturb_gpu_pass_summary));
    const size_t total = o_pass + (out_dev ? 0 : pairs * max_passes * 2 * sizeof(int32_t));
    st = ensure_scratch(b, total);
    if (st != PERTURB_GPU_OK) return st;
    char *base = static_cast<char *>(b->d_summary);
    PB_CUDA(cudaMemcpyAsync(base + o_st, packed.data(), packed.size() * sizeof(double),
                            cudaMemcpyHostToDevice, stream));  // (pageable source: staged before return)
    if (out_dev) {
        st = peer_ready(b, pointer_device(out));
        if (st != PERTURB_GPU_OK) return st;
    }
    pb::launch_gmst(d_ta, d_tb, static_cast<int>(m), reinterpret_cast<double *>(base + o_gm), stream);
    PB_CUDA(cudaGetLastError());

    pb::PropLaunch l;
    fill_launch(b, l);
    l.use_jd = 1;
    l.ta = d_ta;
    l.tb = d_tb;
    l.m = static_cast<int>(m);
    l.n = b->n;
    l.stations = reinterpret_cast<const double *>(base + o_st);
    l.n_stations = static_cast<int>(n_stations);
    l.gmst = reinterpret_cast<const double *>(base + o_gm);
    l.vis_words = reinterpret_cast<uint32_t *>(base + o_words);
    l.vis_best = reinterpret_cast<double2 *>(base + o_best);
    l.pass_out = out_dev ? static_cast<void *>(out) : static_cast<void *>(base + o_out);
    l.passes = (passes == nullptr) ? nullptr
               : out_dev           ? passes
                                   : reinterpret_cast<int32_t *>(base + o_pass);
    l.max_passes = static_cast<int>(max_passes);
    const int launched = pb::launch_sgp4_propagate(l, stream);
    if (launched < 0) {
        return fail(PERTURB_GPU_ERR_ARG, "satellites x times too large for one call: chunk the time grid");
    }
    b->last_launches = launched + 1;
    PB_CUDA(cudaGetLastError());
    if (!out_dev) {
        PB_CUDA(cudaMemcpyAsync(out, base + o_out, pairs * sizeof(perturb_gpu_pass_summary),
                                cudaMemcpyDeviceToHost, stream));
        if (passes != nullptr) {
            PB_CUDA(cudaMemcpyAsync(passes, base + o_pass, pairs * max_passes * 2 * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, stream));
        }
    }
    PB_CUDA(cudaEventRecord(b->ev_call, stream));
    return PERTURB_GPU_OK;
}

int perturb_gpu_propagate_proximity(perturb_gpu_batch *b, const double *jd, const double *jd_frac,
                                    size_t m, const double *primaries, size_t n_primaries,
                                    double threshold_km, perturb_gpu_approach *out,
                                    void *stream_arg) {
    static_assert(sizeof(perturb_gpu_approach) == 16, "approach record is 16 bytes");
    if (b == nullptr) return fail(PERTURB_GPU_ERR_ARG, "batch is null");
    b->last_launches = 0;
    if (b->n == 0 || n_primaries == 0) return PERTURB_GPU_OK;
    if (jd == nullptr || primaries == nullptr || out == nullptr || m == 0 ||
        m > (static_cast<size_t>(1) << 30) || n_primaries > 64 || !(threshold_km >= 0.0)) {
        return fail(PERTURB_GPU_ERR_ARG, "proximity: null pointer, empty / oversized grid, more than "
                                         "64 primaries or a negative threshold");
    }
    const bool out_dev = is_device_pointer(out);
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_arg);
    b->last_stream = stream;
    PB_CUDA(cudaStreamWaitEvent(stream, b->ev_call, 0));
    const double *d_ta, *d_tb;
    int st = stage_times(b, jd, jd_frac, m, stream, &d_ta, &d_tb);
    if (st != PERTURB_GPU_OK) return st;

    const bool prim_dev = is_local_device_pointer(primaries, b->device);
    const size_t chunks = pb::summary_chunks(m), pairs = b->n * n_primaries;
    const size_t prim_bytes = n_primaries * 3 * m * sizeof(double);
    const size_t o_prim = 0, o_part = align256(prim_dev ? 0 : prim_bytes);
    const size_t o_out = align256(o_part + pairs * chunks * pb::prox_partial_bytes());
    const size_t total = o_out + (out_dev ? 0 : pairs * sizeof(perturb_gpu_approach));
    st = ensure_scratch(b, total);
    if (st != PERTURB_GPU_OK) return st;
    char *base = static_cast<char *>(b->d_summary);
    const double *d_prim = primaries;
    if (!prim_dev) {
        PB_CUDA(cudaMemcpyAsync(base + o_prim, primaries, prim_bytes, cudaMemcpyDefault, stream));
        d_prim = reinterpret_cast<const double *>(base + o_prim);
    }
    if (out_dev) {
        st = peer_ready(b, pointer_device(out));
        if (st != PERTURB_GPU_OK) return st;
    }
    pb::PropLaunch l;
    fill_launch(b, l);
    l.use_jd = jd_frac != nullptr ? 1 : 0;
    l.ta = d_ta;
    l.tb = d_tb;
    l.m = static_cast<int>(m);
    l.n = b->n;
    l.primaries = d_prim;
    l.n_primaries = static_cast<int>(n_primaries);
    l.threshold_km = threshold_km;
    l.prox_partials = base + o_part;
    l.approach_out = out_dev ? static_cast<void *>(out) : static_cast<void *>(base + o_out);
    const int launched = pb::launch_sgp4_propagate(l, stream);
    if (launched < 0) {
        return fail(PERTURB_GPU_ERR_ARG, "satellites x times too large for one call: chunk the time grid");
    }
    b->last_launches = launched;
    PB_CUDA(cudaGetLastError());
    if (!out_dev) {
        PB_CUDA(cudaMemcpyAsync(out, base + o_out, pairs * sizeof(perturb_gpu_approach),
                                cudaMemcpyDeviceToHost, stream));
    }
    PB_CUDA(cudaEventRecord(b->ev_call, stream));
    return PERTURB_GPU_OK;
}

int perturb_gpu_synchronize(perturb_gpu_batch *batch) {
    if (batch == nullptr) return fail(PERTURB_GPU_ERR_ARG, "batch is null");
    DeviceGuard guard(batch->device);
    PB_CUDA(cudaStreamSynchronize(batch->last_stream));
    PB_CUDA(cudaStreamSynchronize(batch->copy_stream));
    return PERTURB_GPU_OK;
}

int perturb_gpu_device_count(int *count) {
    if (count == nullptr) return fail(PERTURB_GPU_ERR_ARG, "count is null");
    *count = 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(PERTURB_GPU_ERR_NODEV, "no CUDA device available (there is no CPU fallback)");
    }
    *count = n;
    return PERTURB_GPU_OK;
}

int perturb_gpu_host_alloc(size_t bytes, void **out) {
    if (out == nullptr) return fail(PERTURB_GPU_ERR_ARG, "out is null");
    *out = nullptr;
    if (bytes == 0) return PERTURB_GPU_OK;
    if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(PERTURB_GPU_ERR_ALLOC, "cudaHostAlloc failed");
    }
    return PERTURB_GPU_OK;
}

int perturb_gpu_host_register(void *p, size_t bytes) {
    if (p == nullptr || bytes == 0) return fail(PERTURB_GPU_ERR_ARG, "null pointer or empty range");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return PERTURB_GPU_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PERTURB_GPU_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
    }
    return PERTURB_GPU_OK;
}

void perturb_gpu_host_unregister(void *p) {
    if (p != nullptr && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
}

void perturb_gpu_host_free(void *p) {
    if (p != nullptr) cudaFreeHost(p);
}

int perturb_gpu_device_alloc(int device, size_t bytes, void **out) {
    if (out == nullptr) return fail(PERTURB_GPU_ERR_ARG, "out is null");
    *out = nullptr;
    if (bytes == 0) return PERTURB_GPU_OK;
    DeviceGuard guard(device);
    if (!guard.ok) {
        cudaGetLastError();
        return fail(PERTURB_GPU_ERR_NODEV, "cudaSetDevice failed (there is no CPU fallback)");
    }
    if (cudaMalloc(out, bytes) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(PERTURB_GPU_ERR_ALLOC, "cudaMalloc failed");
    }
    return PERTURB_GPU_OK;
}

void perturb_gpu_device_free(int device, void *p) {
    if (p == nullptr) return;
    DeviceGuard guard(device);
    cudaFree(p);
}

int perturb_gpu_copy(void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return PERTURB_GPU_OK;
    if (dst == nullptr || src == nullptr) return fail(PERTURB_GPU_ERR_ARG, "null pointer");
    PB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return PERTURB_GPU_OK;
}

size_t perturb_gpu_size(const perturb_gpu_batch *batch) { return batch ? batch->n : 0; }

int perturb_gpu_device(const perturb_gpu_batch *batch) { return batch ? batch->device : -1; }

int perturb_gpu_last_error(const perturb_gpu_batch *batch, int32_t *out) {
    if (batch == nullptr || out == nullptr) return fail(PERTURB_GPU_ERR_ARG, "null argument");
    // convert_sgp4_error_code, src/perturb.cpp:24-29
    for (size_t i = 0; i < batch->n; ++i) {
        const int32_t c = batch->init_error[i];
        out[i] = (c < 0 || c >= PERTURB_SGP4_UNKNOWN) ? PERTURB_SGP4_UNKNOWN : c;
    }
    return PERTURB_GPU_OK;
}

int perturb_gpu_orbit_class(const perturb_gpu_batch *batch, uint8_t *out) {
    if (batch == nullptr || out == nullptr) return fail(PERTURB_GPU_ERR_ARG, "null argument");
    if (batch->n) std::memcpy(out, batch->cls.data(), batch->n);
    return PERTURB_GPU_OK;
}

int perturb_gpu_record_field_count(void) { return RF_COUNT; }

const char *perturb_gpu_record_field_name(int field) {
    return (field >= 0 && field < RF_COUNT) ? kRecordFieldNames[field] : nullptr;
}

int perturb_gpu_record_field(const perturb_gpu_batch *batch, int field, double *out) {
    if (batch == nullptr || out == nullptr || field < 0 || field >= RF_COUNT) {
        return fail(PERTURB_GPU_ERR_ARG, "bad argument");
    }
    if (batch->n == 0) return PERTURB_GPU_OK;
    DeviceGuard guard(batch->device);
    PB_CUDA(cudaMemcpy(out, batch->d_rec + static_cast<size_t>(field) * batch->n,
                       batch->n * sizeof(double), cudaMemcpyDeviceToHost));
    return PERTURB_GPU_OK;
}

namespace {

// Members of sgp4::elsetrec by name (generated table, include/perturb_elsetrec_layout.h).
struct Member {
    const char *name;
    size_t offset;
    char kind;  // d double, i int, c char
};
const Member kMembers[] = {
#define X(name, offset, kind) {#name, offset, #kind[0]},
    PERTURB_ELSETREC_MEMBERS(X)
#undef X
};

const Member *find_member(const char *name) {
    for (const Member &m : kMembers) {
        if (std::strcmp(m.name, name) == 0) return &m;
    }
    return nullptr;
}

inline void put_double(void *records, size_t stride, size_t i, const Member *m, double v) {
    std::memcpy(static_cast<char *>(records) + i * stride + m->offset, &v, sizeof(double));
}
inline void put_int(void *records, size_t stride, size_t i, const Member *m, int v) {
    std::memcpy(static_cast<char *>(records) + i * stride + m->offset, &v, sizeof(int));
}
inline void put_char(void *records, size_t stride, size_t i, const Member *m, char v) {
    static_cast<char *>(records)[i * stride + m->offset] = v;
}

}  // namespace

int perturb_gpu_elsetrec_size(void) { return PERTURB_ELSETREC_SIZE; }

int perturb_gpu_export_records(const perturb_gpu_batch *batch, void *records, size_t stride) {
    if (batch == nullptr || records == nullptr || stride < PERTURB_ELSETREC_SIZE) {
        return fail(PERTURB_GPU_ERR_ARG, "null argument or stride below sizeof(sgp4::elsetrec)");
    }
    const size_t n = batch->n;
    if (n == 0) return PERTURB_GPU_OK;
    DeviceGuard guard(batch->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    std::vector<double> plane(n);
    for (int f = 0; f < RF_COUNT; ++f) {
        const char *name = kRecordFieldNames[f];
        const bool is_method = std::strcmp(name, "method_d") == 0;
        const Member *m = find_member(is_method ? "method" : name);
        if (m == nullptr) continue;
        PB_CUDA(cudaMemcpy(plane.data(), batch->d_rec + static_cast<size_t>(f) * n,
                           n * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) {
            if (is_method) {
                put_char(records, stride, i, m, plane[i] != 0.0 ? 'd' : 'n');
            } else if (m->kind == 'd') {
                put_double(records, stride, i, m, plane[i]);
            } else if (m->kind == 'i') {
                put_int(records, stride, i, m, static_cast<int>(plane[i]));
            }
        }
    }
    // what sgp4init leaves besides the numbers: init = 'n' (src/sgp4.cpp:1678), the operation
    // mode it was called with, and t = 0 from its closing sgp4(t = 0) call (:1801)
    const Member *m_init = find_member("init"), *m_ops = find_member("operationmode"),
                 *m_t = find_member("t");
    for (size_t i = 0; i < n; ++i) {
        if (batch->init_error[i] == PERTURB_SGP4_INVALID_TLE) continue;  // never initialised
        put_char(records, stride, i, m_init, 'n');
        put_char(records, stride, i, m_ops, batch->opsmode_a ? 'a' : 'i');
        put_double(records, stride, i, m_t, 0.0);
    }
    return PERTURB_GPU_OK;
}

int perturb_gpu_write_back(perturb_gpu_batch *b, double jd, double jd_frac, int use_jd,
                           void *records, size_t stride) {
    if (b == nullptr || records == nullptr || stride < PERTURB_ELSETREC_SIZE) {
        return fail(PERTURB_GPU_ERR_ARG, "null argument or stride below sizeof(sgp4::elsetrec)");
    }
    const size_t n = b->n;
    if (n == 0) return PERTURB_GPU_OK;
    DeviceGuard guard(b->device);
    if (!guard.ok) return fail(PERTURB_GPU_ERR_CUDA, "cudaSetDevice failed");
    ScratchFree scratch;  // [0] planes, [1] classes
    PB_CUDA(cudaMalloc(&scratch.p[0], static_cast<size_t>(pb::WB_COUNT) * n * sizeof(double)));
    PB_CUDA(cudaMalloc(&scratch.p[1], n));
    PB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_call, 0));
    PB_CUDA(cudaMemcpyAsync(scratch.p[1], b->cls.data(), n, cudaMemcpyHostToDevice, b->stream));
    pb::PropLaunch l;
    fill_launch(b, l);
    l.use_jd = use_jd ? 1 : 0;
    l.n = n;
    pb::launch_write_back(l, static_cast<const uint8_t *>(scratch.p[1]), jd, jd_frac,
                          static_cast<double *>(scratch.p[0]), b->stream);
    PB_CUDA(cudaGetLastError());
    std::vector<double> wb(static_cast<size_t>(pb::WB_COUNT) * n);
    PB_CUDA(cudaMemcpyAsync(wb.data(), scratch.p[0], wb.size() * sizeof(double),
                            cudaMemcpyDeviceToHost, b->stream));
    PB_CUDA(cudaStreamSynchronize(b->stream));
    b->last_launches = 1;

    const Member *m_t = find_member("t"), *m_err = find_member("error");
    const Member *m_mean[7] = {find_member("am"), find_member("em"), find_member("im"),
                               find_member("Om"), find_member("om"), find_member("mm"),
                               find_member("nm")};
    const Member *m_deep[5] = {find_member("aycof"), find_member("xlcof"), find_member("con41"),
                               find_member("x1mth2"), find_member("x7thm1")};
    const Member *m_res[3] = {find_member("atime"), find_member("xli"), find_member("xni")};
    auto at = [&](int plane, size_t i) { return wb[static_cast<size_t>(plane) * n + i]; };
    for (size_t i = 0; i < n; ++i) {
        if (b->init_error[i] == PERTURB_SGP4_INVALID_TLE) {
            // a zeroed record: sgp4() sets t, fails the mean-motion check and stores nothing else
            put_double(records, stride, i, m_t, at(pb::WB_T, i));
            put_int(records, stride, i, m_err, PERTURB_SGP4_MEAN_MOTION);
            continue;
        }
        put_double(records, stride, i, m_t, at(pb::WB_T, i));
        put_int(records, stride, i, m_err, static_cast<int>(at(pb::WB_ERROR, i)));
        const int stage = static_cast<int>(at(pb::WB_STAGE, i));
        const bool deep = b->cls[i] >= PB_CLS_DEEP_NR;
        if (stage >= 1) {
            for (int e = 0; e < 7; ++e) put_double(records, stride, i, m_mean[e], at(pb::WB_AM + e, i));
        }
        if (deep && stage >= 2) {
            put_double(records, stride, i, m_deep[0], at(pb::WB_AYCOF, i));
            put_double(records, stride, i, m_deep[1], at(pb::WB_XLCOF, i));
        }
        if (deep && stage >= 3) {
            put_double(records, stride, i, m_deep[2], at(pb::WB_CON41, i));
            put_double(records, stride, i, m_deep[3], at(pb::WB_X1MTH2, i));
            put_double(records, stride, i, m_deep[4], at(pb::WB_X7THM1, i));
        }
        if (at(pb::WB_RESONANT, i) != 0.0) {
            put_double(records, stride, i, m_res[0], at(pb::WB_ATIME, i));
            put_double(records, stride, i, m_res[1], at(pb::WB_XLI, i));
            put_double(records, stride, i, m_res[2], at(pb::WB_XNI, i));
        }
    }
    return PERTURB_GPU_OK;
}

int perturb_gpu_epoch(const perturb_gpu_batch *batch, double *jd, double *jd_frac) {
    int st = perturb_gpu_record_field(batch, RF_jdsatepoch, jd);
    if (st != PERTURB_GPU_OK) return st;
    return perturb_gpu_record_field(batch, RF_jdsatepochF, jd_frac);
}

int perturb_gpu_last_launch_count(const perturb_gpu_batch *batch) {
    return batch ? batch->last_launches : 0;
}

const char *perturb_gpu_strerror(void) { return g_last_error.c_str(); }

}  // extern "C"
