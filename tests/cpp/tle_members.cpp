// perturb::TwoLineElement::parse as the stand-alone header provides it, member for member:
// reads pairs of lines from stdin and prints every member with 17 significant digits. The
// test feeds the verification catalogue, the ISS example and synthetic TLEs and compares the
// text with the same program built against the reference's own headers and sources.
#include <cstdio>
#include <iostream>
#include <string>

#ifdef PERTURB_BATCH_USE_PERTURB_HPP
#  include <perturb/perturb.hpp>
#else
#  include <perturb/batch_types.hpp>
#endif

using namespace perturb;

int main() {
    std::string l1, l2;
    while (std::getline(std::cin, l1) && std::getline(std::cin, l2)) {
        TwoLineElement t = TwoLineElement();
        l1.resize(TLE_LINE_LEN, ' ');  // verification lines carry a time grid beyond column 69
        l2.resize(TLE_LINE_LEN, ' ');
        const TLEParseError e = t.parse(l1, l2);
        std::printf("err %d\n", static_cast<int>(e));
        if (e != TLEParseError::NONE) {
            continue;
        }
        std::printf("%s %c %u %u %s %u %.17g %.17g %.17g %.17g %u %u %u\n", t.catalog_number,
                    t.classification, t.launch_year, t.launch_number, t.launch_piece, t.epoch_year,
                    t.epoch_day_of_year, t.n_dot, t.n_ddot, t.b_star,
                    static_cast<unsigned>(t.ephemeris_type), t.element_set_number,
                    static_cast<unsigned>(t.line_1_checksum));
        std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %lu %u\n", t.inclination, t.raan,
                    t.eccentricity, t.arg_of_perigee, t.mean_anomaly, t.mean_motion,
                    t.revolution_number, static_cast<unsigned>(t.line_2_checksum));
    }
    return 0;
}
