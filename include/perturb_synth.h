/*
 * perturb_synth.h -- deterministic synthetic TLE catalogues for the benchmark configurations
 * of BASELINE.json. Workload tooling, not part of the drop-in ABI.
 * The generator emits 69-column TLE text with valid checksums, so the SAME bytes feed the
 * reference (`Satellite::from_tle`), the oracle and the GPU batch.
 */
#ifndef PERTURB_SYNTH_H
#define PERTURB_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PERTURB_SYNTH_LEO = 2,        /* C2: near-Earth LEO catalogue (30k x 1440 one-minute steps) */
    PERTURB_SYNTH_MIXED = 3,      /* C3: 70% LEO, 10% MEO, 10% GEO, 5% Molniya, 5% GTO; shuffled */
    PERTURB_SYNTH_DEEP = 4,       /* C4: 50% GEO (irez 1) + 50% Molniya (irez 2)                 */
    PERTURB_SYNTH_MEGA = 5        /* C5: mega-constellation shells                              */
};

#define PERTURB_SYNTH_LINE_STRIDE 70 /* 69 columns + NUL */

/* Default seeds: 0xB2000002 ... 0xB2000005 for configs 2..5 (pass seed = 0 to use them).
 * lines1 / lines2: n records of PERTURB_SYNTH_LINE_STRIDE bytes. Returns 0, or -1 on bad args. */
int perturb_synth_catalog(int config, size_t n, uint64_t seed, char *lines1, char *lines2);

/* Time grid of a config: JulianDate(base_jd, base_frac + k * step_days), k = 0..m-1, where
 * base = catalogue epoch base (2026 day 274.0) + 3 days and step = one minute. */
int perturb_synth_time_grid(size_t m, double *jd, double *jd_frac);

#ifdef __cplusplus
}
#endif
#endif
