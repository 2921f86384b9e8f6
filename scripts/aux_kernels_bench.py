#!/usr/bin/env python
"""Device-timed figures of the kernels either side of the hot path: the batched
ClassicalOrbitalElements (perturb_gpu_rv2coe, 136 B per state) over a config-2 propagate result,
and batch construction from TLE text (device parser + kernel 1 + sort) for config 5's shard.
Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import perturb_b200 as pb
    from perturb_b200 import capi
    dev = torch.device("cuda", 0)
    n, m = 30000, 1440
    b1, b2 = pb.synth_catalog(2, n)
    batch = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", device=0, n=n)
    jd, fr = pb.synth_time_grid(m)
    d_jd, d_fr = torch.tensor(jd, device=dev), torch.tensor(fr, device=dev)
    out = torch.empty((6, n, m), dtype=torch.float64, device=dev)
    err = torch.empty((n, m), dtype=torch.int32, device=dev)
    batch.propagate(d_jd, d_fr, out, err)
    coe = torch.empty((11, n, m), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = capi.lib()
    s = torch.cuda.current_stream(dev)
    times = []
    for it in range(8):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        capi.check(L.perturb_gpu_rv2coe(out.data_ptr(), n, m, m, n * m, int(pb.WGS72), coe.data_ptr(),
                                        n * m, s.cuda_stream), "perturb_gpu_rv2coe")
        z.record(s)
        torch.cuda.synchronize(dev)
        if it >= 3:
            times.append(a.elapsed_time(z))
    ms = sum(times) / len(times)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    gbs = 136.0 * n * m / (ms * 1e-3) * 1e-9
    # construction: 125 000 TLEs (config 5's per-GPU shard) from text
    nb = 125000
    c1, c2 = pb.synth_catalog(5, nb)
    pb.SatelliteBatch.from_text(c1, c2, pb.WGS72, "i", device=0, n=nb).close()
    t0 = time.perf_counter()
    b5 = pb.SatelliteBatch.from_text(c1, c2, pb.WGS72, "i", device=0, n=nb)
    t_text = time.perf_counter() - t0
    b5.close()
    print(json.dumps({
        "rv2coe": {"states": n * m, "ms": ms, "states_per_s": n * m / (ms * 1e-3),
                   "algorithmic_bytes_per_state": 136, "achieved_GBps": gbs, "hbm_peak_GBps": hbm,
                   "frac_of_hbm": gbs / hbm},
        "init_from_text": {"satellites": nb, "wall_ms": 1e3 * t_text,
                           "satellites_per_s": nb / t_text,
                           "what": "H2D of the text, device TLE parser, kernel 1, class sort, pack"},
    }))


if __name__ == "__main__":
    main()
