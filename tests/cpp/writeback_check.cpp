// TEST PROGRAM (integration mode): perturb's own types, the UNMODIFIED reference (oracle/_ref) as
// the checker. Builds `perturb::Satellite`s with the reference, writes the batch's records and
// per-call side effects back into copies of them (SatelliteBatch::export_records / write_back,
// SURVEY.md 8f N2) and compares `sat_rec` member by member with what the reference itself holds
// after the same call sequence:
//   * integer / character members and `t`, `atime`: bit for bit;
//   * double members: difference reported as a fraction of rel * |x| + abs (device math is not
//     glibc's): 1e-11 / 1e-14 for the initialised record, 1e-9 / 1e-10 for the per-call members.
// Compiled here against /root/reference/include by __graft_entry__.build() (the GPU box has no
// reference tree: the binary travels); run by tests/test_cpp_api.py under -m gpu.
//
//     writeback_check <synthetic config> <satellites> [device]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <perturb/perturb.hpp>

#include <perturb/satellite_batch.hpp>
#include <perturb_synth.h>

using namespace perturb;

struct Worst {
    double rel, want, got;
    std::string member;
    std::size_t sat;
    std::size_t exact, doubles, int_mismatch;
    Worst() : rel(0.0), want(0.0), got(0.0), sat(0), exact(0), doubles(0), int_mismatch(0) {}
};

// members the reference never sets from a TLE or leaves to the caller are not compared
static bool compared(const char *name) {
    static const char *skip[] = {"epochtynumrev", "rcse"};
    for (const char *s : skip) {
        if (std::strcmp(s, name) == 0) return false;
    }
    return true;
}

// Difference of one double member as a fraction of its allowance (> 1 fails): `t` and `atime`
// must be the reference's bits (sums of exactly representable terms); everything else may differ
// by rel * |x| + abs -- kernel 1 rounds like the reference apart from libdevice vs glibc (a few
// 1e-14, amplified by 1 - cos^2 style cancellations near i = 0: hence the absolute floor, the same
// allowance as tests/test_gpu_parity.py::test_verification_catalogue_init); the per-call members
// come from kernel 2's fused arithmetic and carry the states' parity bar instead.
static double g_rel = 1e-11, g_abs = 1e-14;

static double metric(const char *name, double x, double y) {
    if (!std::strcmp(name, "t") || !std::strcmp(name, "atime")) return 2.0;
    if (!(x == x) || !(y == y)) return (x != x && y != y) ? 0.0 : 2.0;
    return std::fabs(x - y) / (g_rel * std::fabs(x) + g_abs);
}

static void compare(const std::vector<Satellite> &want, const std::vector<Satellite> &got,
                    const std::vector<bool> &valid, Worst &w) {
    for (std::size_t i = 0; i < want.size(); ++i) {
        if (!valid[i]) continue;
        const char *a = reinterpret_cast<const char *>(&want[i].sat_rec);
        const char *b = reinterpret_cast<const char *>(&got[i].sat_rec);
#define CHECK_MEMBER(name, offset, kind)                                                        \
    if (compared(#name)) {                                                                      \
        if (#kind[0] == 'd') {                                                                  \
            double x, y;                                                                        \
            std::memcpy(&x, a + offset, 8);                                                     \
            std::memcpy(&y, b + offset, 8);                                                     \
            ++w.doubles;                                                                        \
            if (std::memcmp(&x, &y, 8) == 0) {                                                  \
                ++w.exact;                                                                      \
            } else {                                                                            \
                const double d = metric(#name, x, y);                                           \
                if (d > w.rel) {                                                                \
                    w.rel = d;                                                                  \
                    w.member = #name;                                                           \
                    w.want = x;                                                                 \
                    w.got = y;                                                                  \
                    w.sat = i;                                                                  \
                }                                                                               \
            }                                                                                   \
        } else if (#kind[0] == 'i') {                                                           \
            if (std::memcmp(a + offset, b + offset, sizeof(int)) != 0) {                        \
                ++w.int_mismatch;                                                               \
                std::fprintf(stderr, "sat %zu: int member %s differs\n", i, #name);             \
            }                                                                                   \
        } else if (a[offset] != b[offset]) {                                                    \
            ++w.int_mismatch;                                                                   \
            std::fprintf(stderr, "sat %zu: char member %s differs (%d vs %d)\n", i, #name,      \
                         a[offset], b[offset]);                                                 \
        }                                                                                       \
    }
        PERTURB_ELSETREC_MEMBERS(CHECK_MEMBER)
#undef CHECK_MEMBER
    }
}

int main(int argc, char **argv) {
    const int config = argc > 1 ? std::atoi(argv[1]) : 3;
    const std::size_t n = argc > 2 ? static_cast<std::size_t>(std::atoll(argv[2])) : 2000;
    const int device = argc > 3 ? std::atoi(argv[3]) : 0;

    std::vector<char> b1(n * PERTURB_SYNTH_LINE_STRIDE), b2(n * PERTURB_SYNTH_LINE_STRIDE);
    if (perturb_synth_catalog(config, n, 0, b1.data(), b2.data()) != 0) return 1;
    std::vector<std::string> l1(n), l2(n);
    std::vector<Satellite> sats;
    sats.reserve(n);
    std::vector<bool> valid(n, true);
    for (std::size_t i = 0; i < n; ++i) {
        l1[i].assign(&b1[i * PERTURB_SYNTH_LINE_STRIDE], 69);
        l2[i].assign(&b2[i * PERTURB_SYNTH_LINE_STRIDE], 69);
        // the reference: Satellite::from_tle(std::string &, std::string &) hands &line[0] to
        // twoline2rv, which EDITS the text in place (moves the sign of ndot / bstar one column to
        // the left and writes the implied decimal points, src/sgp4.cpp twoline2rv): it gets its own
        // copies, the batch below reads the pristine lines
        std::string c1 = l1[i], c2 = l2[i];
        sats.push_back(Satellite::from_tle(c1, c2));
    }
    SatelliteBatch batch = SatelliteBatch::from_tles(l1.data(), l2.data(), n, GravModel::WGS72, device);
    if (!batch.ok()) {
        std::printf("{\"error\": \"%s\"}\n", SatelliteBatch::status_message());
        return 2;
    }
    std::size_t init_mismatch = 0;
    for (std::size_t i = 0; i < n; ++i) {
        init_mismatch += batch.last_error(i) != sats[i].last_error();
    }

    // (1) the initialised records
    Worst w_init;
    std::vector<Satellite> copy = sats;
    for (Satellite &s : copy) {  // wipe what export_records must restore, keep identity members
        sgp4::elsetrec &r = s.sat_rec;
        sgp4::elsetrec blank = sgp4::elsetrec();
        std::memcpy(blank.satnum, r.satnum, sizeof(r.satnum));
        blank.classification = r.classification;
        std::memcpy(blank.intldesg, r.intldesg, sizeof(r.intldesg));
        blank.epochyr = r.epochyr;
        blank.epochtynumrev = r.epochtynumrev;
        blank.epochdays = r.epochdays;
        blank.ephtype = r.ephtype;
        blank.elnum = r.elnum;
        blank.revnum = r.revnum;
        blank.rcse = r.rcse;
        r = blank;
    }
    if (batch.export_records(copy.data()) != 0) return 3;
    compare(sats, copy, valid, w_init);

    // (2) the per-call side effects, after the same call sequence
    g_rel = 1e-9;   // 1e-6 km over 1e3 km: the states' bar, relative
    g_abs = 1e-10;  // angles [rad], Earth radii
    Worst w_call;
    const double days[] = {0.0, 0.25, 0.5, 3.3, 6.9, -0.35, 1.0};
    std::size_t code_mismatch = 0, calls = 0;
    for (double d : days) {
        std::vector<Sgp4Error> codes(n);
        const JulianDate t = sats[0].epoch() + d;  // one JulianDate for the whole catalogue
        for (std::size_t i = 0; i < n; ++i) {
            StateVector sv;
            codes[i] = sats[i].propagate(t, sv);  // the reference mutates sat_rec
        }
        if (batch.write_back(copy.data(), t) != 0) return 4;
        for (std::size_t i = 0; i < n; ++i) {
            code_mismatch += static_cast<int>(codes[i]) != copy[i].sat_rec.error &&
                             !(codes[i] == Sgp4Error::UNKNOWN);
        }
        compare(sats, copy, valid, w_call);
        ++calls;
    }

    std::printf(
        "{\"satellites\": %zu, \"init_error_mismatch\": %zu, "
        "\"export\": {\"int_char_mismatch\": %zu, \"doubles\": %zu, \"bit_equal\": %zu, "
        "\"worst_ratio\": %.3g, \"worst_member\": \"%s\", \"worst_sat\": %zu, \"want\": %.17g, \"got\": %.17g}, "
        "\"write_back\": {\"calls\": %zu, \"code_mismatch\": %zu, \"int_char_mismatch\": %zu, "
        "\"doubles\": %zu, \"bit_equal\": %zu, \"worst_ratio\": %.3g, \"worst_member\": \"%s\", "
        "\"worst_sat\": %zu}}\n",
        n, init_mismatch, w_init.int_mismatch, w_init.doubles, w_init.exact, w_init.rel,
        w_init.member.c_str(), w_init.sat, w_init.want, w_init.got, calls, code_mismatch, w_call.int_mismatch, w_call.doubles,
        w_call.exact, w_call.rel, w_call.member.c_str(), w_call.sat
    );
    return 0;
}
