// perturb::ShardedSatelliteBatch against one perturb::SatelliteBatch on the same catalogue:
// the sharded result, gathered into one pinned host buffer, must be the unsharded one bit for
// bit (planes and StateVector records), with caller indexing, init errors and epochs intact.
// Shards go to distinct devices when the box has several, else all to device 0.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <perturb/satellite_batch.hpp>
#include <perturb_synth.h>

using namespace perturb;

int main(int argc, char **argv) {
    const std::size_t n = 3000, m = 130;
    const std::size_t n_shards = argc > 1 ? static_cast<std::size_t>(std::atoi(argv[1])) : 3;
    std::vector<char> l1(n * PERTURB_SYNTH_LINE_STRIDE), l2(n * PERTURB_SYNTH_LINE_STRIDE);
    perturb_synth_catalog(PERTURB_SYNTH_MIXED, n, 0, l1.data(), l2.data());
    std::vector<perturb_gpu_tle> parsed(n);
    if (perturb_gpu_parse_tles(l1.data(), l2.data(), PERTURB_SYNTH_LINE_STRIDE, n, 1,
                               parsed.data(), nullptr) != 0) {
        std::printf("FAIL parse\n");
        return 2;
    }
    std::vector<TwoLineElement> tles(n);
    for (std::size_t i = 0; i < n; ++i) {
        tles[i].epoch_year = static_cast<unsigned int>(parsed[i].epoch_year);
        tles[i].epoch_day_of_year = parsed[i].epoch_day_of_year;
        tles[i].n_dot = parsed[i].n_dot;
        tles[i].n_ddot = parsed[i].n_ddot;
        tles[i].b_star = parsed[i].b_star;
        tles[i].inclination = parsed[i].inclination;
        tles[i].raan = parsed[i].raan;
        tles[i].eccentricity = parsed[i].eccentricity;
        tles[i].arg_of_perigee = parsed[i].arg_of_perigee;
        tles[i].mean_anomaly = parsed[i].mean_anomaly;
        tles[i].mean_motion = parsed[i].mean_motion;
    }
    const int n_dev = ShardedSatelliteBatch::device_count();
    if (n_dev <= 0) {
        std::printf("FAIL no device: %s\n", SatelliteBatch::status_message());
        return 2;
    }
    std::vector<int> devices(n_shards);
    for (std::size_t s = 0; s < n_shards; ++s) {
        devices[s] = static_cast<int>(s % static_cast<std::size_t>(n_dev));
    }
    SatelliteBatch whole(tles.data(), n);
    ShardedSatelliteBatch sharded(tles.data(), n, devices.data(), n_shards);
    if (!whole.ok() || !sharded.ok()) {
        std::printf("FAIL init: %s\n", SatelliteBatch::status_message());
        return 3;
    }
    std::printf("devices %d shards %zu cuts", n_dev, sharded.shard_count());
    for (std::size_t s = 0; s < sharded.shard_count(); ++s) {
        std::printf(" %zu", sharded.shard_end(s));
    }
    std::printf("\n");
    int bad = 0;
    for (std::size_t i = 0; i < n; ++i) {
        bad += sharded.last_error(i) != whole.last_error(i);
        bad += sharded.epoch(i).jd != whole.epoch(i).jd;
        bad += sharded.epoch(i).jd_frac != whole.epoch(i).jd_frac;
    }
    std::printf("init_mismatches %d\n", bad);

    std::vector<double> jd(m), fr(m);
    perturb_synth_time_grid(m, jd.data(), fr.data());
    const std::size_t plane = n * m;
    double *pa = static_cast<double *>(ShardedSatelliteBatch::pinned_alloc(6 * plane * 8));
    double *pb = static_cast<double *>(ShardedSatelliteBatch::pinned_alloc(6 * plane * 8));
    std::int32_t *ea = static_cast<std::int32_t *>(ShardedSatelliteBatch::pinned_alloc(plane * 4));
    std::int32_t *eb = static_cast<std::int32_t *>(ShardedSatelliteBatch::pinned_alloc(plane * 4));
    if (!pa || !pb || !ea || !eb) {
        std::printf("FAIL pinned_alloc\n");
        return 4;
    }
    whole.propagate_soa(jd.data(), fr.data(), m, pa, ea, m, plane, nullptr);
    sharded.propagate_soa(jd.data(), fr.data(), m, pb, eb, m, plane);
    std::printf("planes_equal %d codes_equal %d\n",
                std::memcmp(pa, pb, 6 * plane * 8) == 0, std::memcmp(ea, eb, plane * 4) == 0);

    // device-resident gather: the same call with buffers on the LAST device; shards on other
    // devices store their rows there over NVLink peer access
    {
        const int target = n_dev - 1;
        double *dp = static_cast<double *>(ShardedSatelliteBatch::device_alloc(target, 6 * plane * 8));
        std::int32_t *de =
            static_cast<std::int32_t *>(ShardedSatelliteBatch::device_alloc(target, plane * 4));
        std::memset(pb, 0xff, 6 * plane * 8);
        std::memset(eb, 0xff, plane * 4);
        int st = (dp && de) ? sharded.propagate_soa(jd.data(), fr.data(), m, dp, de, m, plane) : -99;
        if (st == 0) st = ShardedSatelliteBatch::copy(pb, dp, 6 * plane * 8);
        if (st == 0) st = ShardedSatelliteBatch::copy(eb, de, plane * 4);
        std::printf("device_gather_equal %d device_gather_codes_equal %d status %d target %d\n",
                    std::memcmp(pa, pb, 6 * plane * 8) == 0, std::memcmp(ea, eb, plane * 4) == 0, st,
                    target);
        ShardedSatelliteBatch::device_free(target, dp);
        ShardedSatelliteBatch::device_free(target, de);
    }

    std::vector<double> mins(m);
    for (std::size_t k = 0; k < m; ++k) {
        mins[k] = 11.0 * static_cast<double>(k) - 100.0;
    }
    StateVector *sa = static_cast<StateVector *>(ShardedSatelliteBatch::pinned_alloc(plane * 64));
    StateVector *sb = static_cast<StateVector *>(ShardedSatelliteBatch::pinned_alloc(plane * 64));
    std::vector<Sgp4Error> ca(plane), cb(plane);
    whole.propagate_states(mins.data(), nullptr, m, sa, ca.data(), m, nullptr);
    sharded.propagate_states(mins.data(), nullptr, m, sb, cb.data(), m);
    std::printf("records_equal %d record_codes_equal %d status %d\n",
                std::memcmp(sa, sb, plane * 64) == 0,
                std::memcmp(ca.data(), cb.data(), plane * 4) == 0, sharded.status());
    // the drop-in calls on the sharded class: ordinary vectors, untouched on errors 1-4, every
    // shard on its own host thread -- and the from_tles route (text parsed on each shard's device)
    {
        std::vector<JulianDate> times(m);
        for (std::size_t k = 0; k < m; ++k) times[k] = JulianDate(jd[k], fr[k]);
        StateVector mark = StateVector();
        mark.position = {{-1.0, -2.0, -3.0}};
        mark.velocity = {{-4.0, -5.0, -6.0}};
        std::vector<StateVector> va(plane, mark), vb(plane, mark);
        std::vector<Sgp4Error> xa(plane), xb(plane);
        whole.propagate(times.data(), m, va.data(), xa.data());
        std::vector<std::string> s1(n), s2(n);
        for (std::size_t i = 0; i < n; ++i) {
            s1[i].assign(&l1[i * PERTURB_SYNTH_LINE_STRIDE], 69);
            s2[i].assign(&l2[i * PERTURB_SYNTH_LINE_STRIDE], 69);
        }
        ShardedSatelliteBatch from_text = ShardedSatelliteBatch::from_tles(
            s1.data(), s2.data(), n, devices.data(), n_shards);
        const int st = from_text.ok() ? from_text.propagate(times.data(), m, vb.data(), xb.data()) : -98;
        std::size_t untouched = 0;
        for (std::size_t e = 0; e < plane; ++e) {
            const int c = static_cast<int>(xa[e]);
            if (c >= 1 && c <= 4) {
                untouched += std::memcmp(&va[e].position, &mark.position, 48) == 0;
            }
        }
        std::size_t failed = 0;
        for (std::size_t e = 0; e < plane; ++e) {
            const int c = static_cast<int>(xa[e]);
            failed += (c >= 1 && c <= 4);
        }
        std::printf("dropin_equal %d dropin_codes_equal %d status %d failed %zu untouched %zu cuts_equal %d\n",
                    std::memcmp(va.data(), vb.data(), plane * 64) == 0,
                    std::memcmp(xa.data(), xb.data(), plane * 4) == 0, st, failed, untouched,
                    from_text.shard_end(0) == sharded.shard_end(0));
    }
    ShardedSatelliteBatch::pinned_free(pa);
    ShardedSatelliteBatch::pinned_free(pb);
    ShardedSatelliteBatch::pinned_free(ea);
    ShardedSatelliteBatch::pinned_free(eb);
    ShardedSatelliteBatch::pinned_free(sa);
    ShardedSatelliteBatch::pinned_free(sb);
    return 0;
}
