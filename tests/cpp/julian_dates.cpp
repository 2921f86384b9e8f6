// perturb::JulianDate / perturb::DateTime as the batch's stand-alone header declares them,
// exercised like the reference's own unit tests (tests/test_perturb.cpp:53-129): conversions
// both ways over 1901-2099, normalisation, offsets and ordering. Prints every result with 17
// significant digits; the test compares the text with the same program built against the
// reference's own headers and sources (-DPERTURB_BATCH_USE_PERTURB_HPP).
#include <cstdint>
#include <cstdio>

#include <perturb/batch_types.hpp>

using namespace perturb;

static std::uint64_t state = 0x9E3779B97F4A7C15ull;
static std::uint64_t next_u64() {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return state;
}
static int pick(int lo, int hi) {  // inclusive
    return lo + static_cast<int>(next_u64() % static_cast<std::uint64_t>(hi - lo + 1));
}

static void show(const char *tag, const JulianDate &j) {
    const DateTime t = j.to_datetime();
    std::printf("%s %.17g %.17g -> %d %d %d %d %d %.17g\n", tag, j.jd, j.jd_frac, t.year, t.month,
                t.day, t.hour, t.min, t.sec);
}

int main() {
    const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    // every first and last day of every month, 1901-2099, at the day's edges
    for (int y = 1901; y <= 2099; ++y) {
        for (int m = 1; m <= 12; ++m) {
            const int last = mdays[m - 1] + ((m == 2 && y % 4 == 0) ? 1 : 0);
            show("edge", JulianDate(DateTime{y, m, 1, 0, 0, 0.0}));
            show("edge", JulianDate(DateTime{y, m, last, 23, 59, 59.999}));
        }
    }
    // random time points
    for (int i = 0; i < 20000; ++i) {
        const int y = pick(1901, 2099), m = pick(1, 12);
        const int last = mdays[m - 1] + ((m == 2 && y % 4 == 0) ? 1 : 0);
        const double sec = static_cast<double>(next_u64() >> 11) * (60.0 / 9007199254740992.0);
        show("rand", JulianDate(DateTime{y, m, pick(1, last), pick(0, 23), pick(0, 59), sec}));
    }
    // arithmetic, normalisation, ordering
    JulianDate a(DateTime{2022, 3, 14, 1, 59, 26.535});
    JulianDate b = a + 3.75;
    show("plus", b);
    show("norm", b.normalized());
    b += 100.125;
    show("pluseq", b);
    b.normalize();
    show("normeq", b);
    JulianDate c = a - 0.5;
    show("minus", c);
    c -= 1000.0625;
    show("minuseq", c);
    show("norm", c.normalized());
    show("norm", JulianDate(2459652.75, 0.5).normalized());
    show("norm", JulianDate(2459652.0, -2.25).normalized());
    show("norm", JulianDate(2459652.123456789).normalized());
    std::printf("diff %.17g %.17g\n", b - a, a - c);
    std::printf("order %d %d %d %d %d %d %d %d\n", a < b, a > b, a <= b, a >= b, a < a, a <= a,
                a >= a, c > a);
    return 0;
}
