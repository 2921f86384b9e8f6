#!/usr/bin/env python
"""Short time grids (a whole catalogue at one epoch, or a few): device-timed evaluations/s of
perturb_gpu_propagate for m = 1 .. 256 on the mixed 100k catalogue. Prints one
JSON line; PERTURB_GPU_LIB selects a library variant."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import perturb_b200 as pb
    dev = torch.device("cuda", 0)
    n = 100000
    b1, b2 = pb.synth_catalog(3, n)
    batch = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", device=0, n=n)
    s = torch.cuda.current_stream(dev)
    res = {}
    for m in (1, 2, 4, 8, 16, 17, 32, 64, 96, 128, 256):
        jd, fr = pb.synth_time_grid(m)
        d_jd, d_fr = torch.tensor(jd, device=dev), torch.tensor(fr, device=dev)
        out = torch.empty((6, n, m), dtype=torch.float64, device=dev)
        err = torch.empty((n, m), dtype=torch.int32, device=dev)
        ts = []
        for it in range(8):
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            batch.propagate(d_jd, d_fr, out, err, stream=s.cuda_stream, sync=False)
            z.record(s)
            torch.cuda.synchronize(dev)
            if it >= 3:
                ts.append(a.elapsed_time(z))
        ms = sum(ts) / len(ts)
        res[str(m)] = {"ms": ms, "evals_per_s": n * m / (ms * 1e-3)}
    print(json.dumps({"lib": os.path.basename(pb.capi.LIB_PATH), "satellites": n,
                      "what": "mixed 100k catalogue, device-resident outputs", "by_times": res}))


if __name__ == "__main__":
    main()
