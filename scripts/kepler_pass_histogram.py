#!/usr/bin/env python
"""Kepler-solve statistics of a synthetic catalogue (CPU only): how many sin/cos passes the
reference's Newton iteration (src/sgp4.cpp:1965-1983) takes per evaluation, how many the
round-1 device loop took (it stopped one pass earlier and rotated), and the share of
evaluations that the closed-form path of round 2 (`kepler_small`, e^2 < 4e-4: ONE general sin/cos)
covers. The eccentricity is the TLE's (the long-period term aynl adds ~1e-3 to it); the argument
of latitude is sampled uniformly.

    python scripts/kepler_pass_histogram.py [config=2] [satellites=30000] > profiles/r2_kepler_passes_c2.json
"""
import json
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import perturb_b200 as pb  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
b1, b2 = pb.synth_catalog(cfg, n)
L2 = pb.block_to_lines(b2, n)
ecc = np.array([float("0." + ln[26:33]) for ln in L2])
rng = np.random.default_rng(7)
ref_passes = np.zeros(12, dtype=np.int64)
r1_passes = np.zeros(12, dtype=np.int64)
samples = 8
for e in ecc:
    for _ in range(samples):
        ph, u = rng.uniform(0, 2 * math.pi, 2)
        axnl, aynl = e * math.cos(ph), e * math.sin(ph)
        eo1, tem5, ktr, owed = u, 9999.9, 1, None
        while abs(tem5) >= 1e-12 and ktr <= 10:
            s, c = math.sin(eo1), math.cos(eo1)
            tem5 = (u - aynl * c + axnl * s - eo1) / (1.0 - c * axnl - s * aynl)
            tem5 = max(-0.95, min(0.95, tem5))
            if owed is None and abs(tem5) < 5.9604644775390625e-8:
                owed = ktr  # round 1 stopped here and rotated by the correction
            eo1 += tem5
            ktr += 1
        ref_passes[ktr - 1] += 1
        r1_passes[owed if owed is not None else ktr - 1] += 1
tot = float(ref_passes.sum())
out = {
    "config": cfg, "satellites": n, "samples_per_satellite": samples,
    "eccentricity_quantiles": {q: float(np.quantile(ecc, float(q))) for q in ("0.5", "0.9", "0.99", "1.0")},
    "share_closed_form_e2_below_4e-4": float((ecc * ecc < 4e-4).mean()),
    "sincos_passes_reference_loop": {str(k): ref_passes[k] / tot for k in range(12) if ref_passes[k]},
    "sincos_passes_round1_device_loop": {str(k): r1_passes[k] / tot for k in range(12) if r1_passes[k]},
    "mean_passes_reference": float((np.arange(12) * ref_passes).sum() / tot),
    "mean_passes_round1": float((np.arange(12) * r1_passes).sum() / tot),
    "mean_general_sincos_round2": float((ecc * ecc < 4e-4).mean() * 1.0
                                        + (ecc * ecc >= 4e-4).mean() * (np.arange(12) * r1_passes).sum() / tot),
}
print(json.dumps(out, indent=1))
