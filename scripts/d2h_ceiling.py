#!/usr/bin/env python
"""Raw device -> pinned-host ceiling of the box, one process driving 1, 2, 4, ... all GPUs at once,
for the pinned-allocation flavours (default, portable, write-combined): separates "the box cannot
land more bytes in host memory" from "the buffers were placed badly". One JSON line per case.
    python scripts/d2h_ceiling.py [GiB per copy=1] [reps=4]"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
from perturb_b200 import capi  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = capi.tools()
import torch  # noqa: E402  (device count only)

ndev = torch.cuda.device_count()
counts = [k for k in (1, 2, 4, 8) if k <= ndev]
for flags, name in ((0, "default"), (1, "portable"), (4, "write-combined")):
    for k in counts:
        g = C.c_double(0.0)
        st = T.perturb_gpu_measure_d2h_multi(k, int(gib * (1 << 30)), reps, flags, C.byref(g))
        print(json.dumps({"one_process_devices": k, "pinned": name, "status": st,
                          "aggregate_d2h_gbs": g.value, "per_device_gbs": g.value / k}), flush=True)
