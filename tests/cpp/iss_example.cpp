// The reference's own example (examples/cmake-local/main.cpp:8-33) and a slice of its
// verification test, rewritten against perturb::SatelliteBatch. Prints one line per check;
// tests/test_cpp_api.py compares the numbers with the golden fixtures.
#include <cstdio>
#include <string>
#include <vector>

#include <perturb/satellite_batch.hpp>

using namespace perturb;

int main() {
    std::vector<std::string> l1 = {
        "1 25544U 98067A   22071.78032407  .00021395  00000-0  39008-3 0  9996",
        "1 00005U 58002B   00179.78495062  .00000023  00000-0  28098-4 0  4753",
        "1 08195U 75081A   06176.33215444  .00000099  00000-0  11873-3 0   813",
        "1 25544U 98067A   22071.7803"};  // too short: INVALID_TLE
    std::vector<std::string> l2 = {
        "2 25544  51.6424  94.0370 0004047 256.5103  89.8846 15.49386383330227",
        "2 00005  34.2682 348.7242 1859667 331.7664  19.3264 10.82419157413667",
        "2 08195  64.1586 279.0717 6877146 264.7651  20.2257  2.00491383225656",
        "2 25544  51.6424  94.0370 0004047 256.5103  89.8846 15.49386383330227"};

    auto batch = SatelliteBatch::from_tles(l1.data(), l2.data(), l1.size());
    if (!batch.ok()) {
        std::printf("FAIL status %d: %s\n", batch.status(), SatelliteBatch::status_message());
        return 2;
    }
    for (std::size_t i = 0; i < batch.size(); ++i) {
        std::printf("init %zu err %d epoch %.17g %.17g\n", i,
                    static_cast<int>(batch.last_error(i)), batch.epoch(i).jd,
                    batch.epoch(i).jd_frac);
    }

    // as in the reference's example (examples/cmake-local/main.cpp:21)
    const std::vector<JulianDate> times = {JulianDate(DateTime{2022, 3, 14, 1, 59, 26.535})};
    std::vector<StateVector> svs(batch.size() * times.size());
    for (auto &sv : svs) {
        sv.position = {{-1.0, -1.0, -1.0}};  // must stay untouched on errors 1-4
        sv.velocity = {{-1.0, -1.0, -1.0}};
    }
    std::vector<Sgp4Error> errs(svs.size());
    if (batch.propagate(times.data(), times.size(), svs.data(), errs.data()) != 0) {
        std::printf("FAIL propagate: %s\n", SatelliteBatch::status_message());
        return 3;
    }
    for (std::size_t i = 0; i < batch.size(); ++i) {
        const auto &sv = svs[i];
        std::printf("jd %zu err %d r %.17g %.17g %.17g v %.17g %.17g %.17g\n", i,
                    static_cast<int>(errs[i]), sv.position[0], sv.position[1], sv.position[2],
                    sv.velocity[0], sv.velocity[1], sv.velocity[2]);
        std::printf("jdepoch %zu %.17g %.17g\n", i, sv.epoch.jd, sv.epoch.jd_frac);
    }

    const std::vector<double> mins = {0.0, 360.0, 720.0};
    std::vector<StateVector> sv2(batch.size() * mins.size());
    std::vector<Sgp4Error> err2(sv2.size());
    batch.propagate_from_epoch(mins.data(), mins.size(), sv2.data(), err2.data());
    for (std::size_t i = 0; i < batch.size(); ++i) {
        for (std::size_t k = 0; k < mins.size(); ++k) {
            const auto &sv = sv2[i * mins.size() + k];
            std::printf("mins %zu %g err %d r %.17g %.17g %.17g v %.17g %.17g %.17g\n", i,
                        mins[k], static_cast<int>(err2[i * mins.size() + k]), sv.position[0],
                        sv.position[1], sv.position[2], sv.velocity[0], sv.velocity[1],
                        sv.velocity[2]);
            // perturb.cpp:229: sv.epoch = epoch() + mins / 1440 (added to the fraction)
            const JulianDate want = batch.epoch(i) + (mins[k] / 1440.0);
            std::printf("minsepoch %zu %g %d\n", i, mins[k],
                        static_cast<int>(sv.epoch.jd == want.jd && sv.epoch.jd_frac == want.jd_frac));
        }
    }
    return 0;
}
