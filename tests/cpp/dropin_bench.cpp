// Times the drop-in call itself, as user code makes it:
//
//     std::vector<StateVector> svs(n * m);  std::vector<Sgp4Error> errs(n * m);
//     batch.propagate(times.data(), m, svs.data(), errs.data());        // perturb.cpp:236-242, batched
//
// i.e. ordinary (pageable) std::vector destinations, the reference's untouched-on-error
// semantics, host time grid in, everything inside the timed region. `bench.py` runs this
// program for its `e2e_dropin` figure; next to it the same call into PINNED records
// (ShardedSatelliteBatch::pinned_alloc + propagate_states) for comparison.
//
//     dropin_bench <synthetic config 2..5> <satellites> <times> <steps> [device]
//
// Prints one JSON line. Needs a GPU (there is no CPU fallback): exit code 2 otherwise.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <perturb/satellite_batch.hpp>
#include <perturb_synth.h>

using namespace perturb;

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s config satellites times steps [device]\n", argv[0]);
        return 1;
    }
    const int config = std::atoi(argv[1]);
    const std::size_t n = static_cast<std::size_t>(std::atoll(argv[2]));
    const std::size_t m = static_cast<std::size_t>(std::atoll(argv[3]));
    const int steps = std::atoi(argv[4]);
    const int device = argc > 5 ? std::atoi(argv[5]) : 0;

    std::vector<char> b1(n * PERTURB_SYNTH_LINE_STRIDE), b2(n * PERTURB_SYNTH_LINE_STRIDE);
    if (perturb_synth_catalog(config, n, 0, b1.data(), b2.data()) != 0) {
        std::fprintf(stderr, "bad synthetic catalogue arguments\n");
        return 1;
    }
    std::vector<std::string> l1(n), l2(n);
    for (std::size_t i = 0; i < n; ++i) {
        l1[i].assign(&b1[i * PERTURB_SYNTH_LINE_STRIDE], 69);
        l2[i].assign(&b2[i * PERTURB_SYNTH_LINE_STRIDE], 69);
    }
    std::vector<double> jd(m), fr(m);
    perturb_synth_time_grid(m, jd.data(), fr.data());
    std::vector<JulianDate> times(m);
    for (std::size_t k = 0; k < m; ++k) times[k] = JulianDate(jd[k], fr[k]);

    const double t_init0 = now_s();
    SatelliteBatch batch = SatelliteBatch::from_tles(l1.data(), l2.data(), n, GravModel::WGS72, device);
    if (!batch.ok()) {
        std::printf("{\"error\": \"%s\"}\n", SatelliteBatch::status_message());
        return 2;
    }
    const double init_s = now_s() - t_init0;

    // ---- the drop-in call into ordinary vectors ----
    std::vector<StateVector> svs(n * m);
    std::vector<Sgp4Error> errs(n * m);
    if (batch.propagate(times.data(), m, svs.data(), errs.data()) != 0) {  // warm-up (buffers)
        std::printf("{\"error\": \"%s\"}\n", SatelliteBatch::status_message());
        return 3;
    }
    const double t0 = now_s();
    for (int s = 0; s < steps; ++s) {
        if (batch.propagate(times.data(), m, svs.data(), errs.data()) != 0) return 3;
    }
    const double dropin_s = (now_s() - t0) / steps;
    std::size_t ok = 0;
    double checksum = 0.0;
    for (std::size_t e = 0; e < n * m; ++e) {
        if (errs[e] == Sgp4Error::NONE) {
            ++ok;
            checksum += svs[e].position[0];
        }
    }

    // ---- the same call after ONE more line of user code: the vectors page-locked in place ----
    double registered_s = 0.0, register_s = 0.0;
    bool registered_same = true;
    {
        std::vector<StateVector> svr(n * m);
        std::vector<Sgp4Error> errr(n * m);
        const double tr0 = now_s();
        PinnedRegion pin_sv(svr.data(), svr.size() * sizeof(StateVector));
        PinnedRegion pin_err(errr.data(), errr.size() * sizeof(Sgp4Error));
        register_s = now_s() - tr0;
        if (pin_sv.ok() && pin_err.ok()) {
            if (batch.propagate(times.data(), m, svr.data(), errr.data()) != 0) return 5;  // warm-up
            const double t2 = now_s();
            for (int s = 0; s < steps; ++s) {
                if (batch.propagate(times.data(), m, svr.data(), errr.data()) != 0) return 5;
            }
            registered_s = (now_s() - t2) / steps;
            registered_same = std::memcmp(svr.data(), svs.data(), svr.size() * sizeof(StateVector)) == 0 &&
                              std::memcmp(errr.data(), errs.data(), errr.size() * sizeof(Sgp4Error)) == 0;
        }
    }

    // ---- the same records into pinned memory (copy engine writes them directly) ----
    double pinned_s = 0.0;
    bool same = true;
    StateVector *psv = static_cast<StateVector *>(
        ShardedSatelliteBatch::pinned_alloc(n * m * sizeof(StateVector)));
    Sgp4Error *perr = static_cast<Sgp4Error *>(
        ShardedSatelliteBatch::pinned_alloc(n * m * sizeof(Sgp4Error)));
    if (psv != nullptr && perr != nullptr) {
        batch.propagate_states(jd.data(), fr.data(), m, psv, perr, m, nullptr);  // warm-up
        const double t1 = now_s();
        for (int s = 0; s < steps; ++s) {
            if (batch.propagate_states(jd.data(), fr.data(), m, psv, perr, m, nullptr) != 0) return 4;
        }
        pinned_s = (now_s() - t1) / steps;
        // both routes carry the same bits wherever the reference writes (codes 0 and 6)
        for (std::size_t e = 0; e < n * m && same; ++e) {
            same = perr[e] == errs[e] &&
                   std::memcmp(&psv[e].epoch, &svs[e].epoch, sizeof(JulianDate)) == 0;
            if (same && (errs[e] == Sgp4Error::NONE || errs[e] == Sgp4Error::DECAYED)) {
                same = std::memcmp(&psv[e], &svs[e], sizeof(StateVector)) == 0;
            }
        }
    }
    ShardedSatelliteBatch::pinned_free(psv);
    ShardedSatelliteBatch::pinned_free(perr);

    const double evals = static_cast<double>(n) * static_cast<double>(m);
    std::printf(
        "{\"config\": %d, \"satellites\": %zu, \"times\": %zu, \"steps\": %d, "
        "\"dropin_evals_per_s\": %.6g, \"dropin_ms\": %.4f, \"pinned_evals_per_s\": %.6g, "
        "\"pinned_ms\": %.4f, \"dropin_equals_pinned\": %s, \"ok_fraction\": %.6f, "
        "\"checksum_x\": %.17g, \"bytes_per_eval\": 68, \"init_s\": %.4f, "
        "\"registered_evals_per_s\": %.6g, \"registered_ms\": %.4f, \"register_once_ms\": %.2f, "
        "\"registered_equals_dropin\": %s}\n",
        config, n, m, steps, evals / dropin_s, dropin_s * 1e3,
        pinned_s > 0.0 ? evals / pinned_s : 0.0, pinned_s * 1e3, same ? "true" : "false",
        static_cast<double>(ok) / evals, checksum, init_s,
        registered_s > 0.0 ? evals / registered_s : 0.0, registered_s * 1e3, register_s * 1e3,
        registered_same ? "true" : "false"
    );
    return 0;
}
