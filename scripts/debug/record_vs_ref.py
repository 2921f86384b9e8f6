#!/usr/bin/env python
"""Kernel 1's record (perturb_gpu_record_field, and the same numbers through
perturb_gpu_export_records) against the compiled reference, field by field, on a synthetic
catalogue: the worst field / satellite and every field of one named satellite.
    python scripts/debug/record_vs_ref.py <config> <n> [satellite]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import perturb_b200 as pb  # noqa: E402
from oracle import ref  # noqa: E402

cfg, n = int(sys.argv[1]), int(sys.argv[2])
one = int(sys.argv[3]) if len(sys.argv) > 3 else None
b1, b2 = pb.synth_catalog(cfg, n)
L1, L2 = pb.block_to_lines(b1, n), pb.block_to_lines(b2, n)
batch = pb.SatelliteBatch.from_tles(L1, L2)
sats = ref.RefSatellites.from_tle_lines(L1, L2)
want = sats.fields()  # [n][fields] in ref.FIELD_NAMES order
names = pb.SatelliteBatch.record_field_names()
got = {nm: batch.record_field(nm) for nm in names}
rows = []
for j, nm in enumerate(ref.FIELD_NAMES):
    if nm not in got:
        continue
    w, g = want[:, j], got[nm]
    d = np.abs(w - g) / (1e-11 * np.abs(w) + 1e-14)
    d[np.isnan(w) & np.isnan(g)] = 0.0
    i = int(np.nanargmax(d))
    rows.append((float(d[i]), nm, i, float(w[i]), float(g[i])))
rows.sort(reverse=True)
for r in rows[:8]:
    print(json.dumps({"ratio": r[0], "field": r[1], "sat": r[2], "ref": r[3], "gpu": r[4]}))
rec = batch.export_records()
layout = dict((ln.split("(")[1].split(",")[0], int(ln.split(",")[1]))
              for ln in open("include/perturb_elsetrec_layout.h") if ln.strip().startswith("X(") and ", d)" in ln)
if one is not None:
    for j, nm in enumerate(ref.FIELD_NAMES):
        if nm in got and nm in layout:
            e = rec[one, layout[nm]:layout[nm] + 8].copy().view(np.float64)[0]
            if want[one, j] != got[nm][one] or e != got[nm][one]:
                print(nm, "ref", want[one, j], "record_field", got[nm][one], "export", e)
