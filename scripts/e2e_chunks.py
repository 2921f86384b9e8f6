#!/usr/bin/env python
"""End-to-end rate of the pinned-host output path against the number of time-chunks the copy is
split into (PERTURB_GPU_HOST_CHUNKS; 1 = one linear copy after the kernel, more = kernel hidden
behind the previous chunk's copy but shorter 2-D copy rows). One JSON line per setting."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import perturb_b200 as pb  # noqa: E402
import torch  # noqa: E402

n, m = 30000, 1440
b1, b2 = pb.synth_catalog(2, n)
batch = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", n=n)
jd, fr = pb.synth_time_grid(m)
h_jd, h_fr = torch.tensor(jd).pin_memory(), torch.tensor(fr).pin_memory()
h_out = torch.empty((6, n, m), dtype=torch.float64).pin_memory()
h_err = torch.empty((n, m), dtype=torch.int32).pin_memory()
h_rec = torch.empty((n, m, 8), dtype=torch.float64).pin_memory()
for chunks in (1, 2, 3, 4, 6, 12):
    os.environ["PERTURB_GPU_HOST_CHUNKS"] = str(chunks)
    out = {"chunks": chunks}
    for name, call, nbytes in (("planes", lambda: batch.propagate(h_jd, h_fr, h_out, h_err), 52),
                               ("records", lambda: batch.propagate_states(h_jd, h_fr, h_rec, h_err), 68)):
        call()
        t0 = time.perf_counter()
        for _ in range(4):
            call()
        s = (time.perf_counter() - t0) / 4
        out[name + "_evals_per_s"] = n * m / s
        out[name + "_gbs"] = n * m * nbytes / s * 1e-9
    print(json.dumps(out), flush=True)
