#!/usr/bin/env python
"""Device-resident gather of a sharded catalogue (opt-in): ONE process, G GPUs,
every shard's kernel 2 storing its satellite rows straight into device 0's result planes over
NVLink peer access (fused compute + transfer, no staging, no collective).

    python scripts/p2p_gather_bench.py [--workload c2] [--steps 10] [--devices 2]

Prints one JSON line: evaluations/s of the whole job (max over devices of the CUDA-event time
of each shard's launches), next to the same shards writing device-local buffers.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import perturb_b200 as pb
    from bench import WORKLOADS

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--devices", type=int, default=0)
    ap.add_argument("--times", type=int, default=0, help="cut of the time grid (0 = all)")
    args = ap.parse_args()
    g = args.devices or torch.cuda.device_count()
    cfg, n_per, m = WORKLOADS[args.workload]
    if args.times:
        m = args.times
    n = n_per * g
    b1, b2 = pb.synth_catalog(cfg, n)
    tles, _ = pb.parse_tle_blocks(b1, b2, n, flavor=1)
    sharded = pb.ShardedSatelliteBatch.from_parsed(tles, list(range(g)))
    jd, fr = pb.synth_time_grid(m)

    def run(make_out):
        outs = make_out()
        grids = [(torch.tensor(jd, device="cuda:%d" % d), torch.tensor(fr, device="cuda:%d" % d))
                 for d in range(g)]
        for d in range(g):
            torch.cuda.synchronize(d)
        evs = []
        total = []
        for it in range(3 + args.steps):
            step = []
            for d, (b, (lo, hi)) in enumerate(zip(sharded.batches, sharded.ranges)):
                with torch.cuda.device(d):
                    s = torch.cuda.current_stream(d)
                    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(s)
                    pv, er = outs(d, lo, hi)
                    b.propagate(grids[d][0], grids[d][1], pv, er, stream=s.cuda_stream, sync=False)
                    z.record(s)
                    step.append((a, z))
            for d in range(g):
                torch.cuda.synchronize(d)
            if it >= 3:
                total.append(max(a.elapsed_time(z) for a, z in step))
        ms = float(np.mean(total))
        return n * m / (ms * 1e-3), ms

    def gathered():
        pv = torch.empty((6, n, m), dtype=torch.float64, device="cuda:0")
        er = torch.empty((n, m), dtype=torch.int32, device="cuda:0")
        return lambda d, lo, hi: (pv[:, lo:hi], er[lo:hi])

    def local():
        bufs = []
        for d, (lo, hi) in enumerate(sharded.ranges):
            bufs.append((torch.empty((6, hi - lo, m), dtype=torch.float64, device="cuda:%d" % d),
                         torch.empty((hi - lo, m), dtype=torch.int32, device="cuda:%d" % d)))
        return lambda d, lo, hi: bufs[d]

    v_local, ms_local = run(local)
    v_gath, ms_gath = run(gathered)
    remote_bytes = 52.0 * (n - (sharded.ranges[0][1] - sharded.ranges[0][0])) * m
    print(json.dumps({
        "what": "one process, %d GPUs, shards of %d satellites x %d times; device-resident gather "
                "on cuda:0 by peer stores from kernel 2 vs device-local outputs" % (g, n_per, m),
        "workload": args.workload, "devices": g, "steps": args.steps,
        "local_evals_per_s": v_local, "local_ms_per_step": ms_local,
        "gathered_evals_per_s": v_gath, "gathered_ms_per_step": ms_gath,
        "nvlink_in_GBps_on_target": remote_bytes / (ms_gath * 1e-3) * 1e-9,
        "remote_bytes_per_step": remote_bytes,
    }))


if __name__ == "__main__":
    main()
