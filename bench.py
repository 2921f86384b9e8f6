#!/usr/bin/env python
"""Benchmark of the hot path: batched SGP4 `Satellite::propagate` on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]
                    [--quick | --e2e-only | --no-extra]

A "step" is one pass of the hot path over one batch: every satellite of the rank's catalogue
shard propagated to every time of the grid (BASELINE.json configs[1]: 30 000 near-Earth
satellites x 1440 one-minute steps per GPU; weak scaling: the catalogue is N x 30 000
satellites sharded by satellite range, no data-path collective).

Prints ONE JSON line (rank 0). Keys beyond the base contract:
  roofline      FP64-pipe roofline of kernel 2 (the binding one): algorithmic
                flops per evaluation (flops-v0, frozen in F_ALG below) x evaluations per
                launch / measured launch time, against a DFMA peak measured in this run.
  roofline_hbm  the second roofline: 52 B written per evaluation against MEASURED_PEAKS.json.
  cpu_baseline  the reference's scalar loop (oracle/_ref, else the C port) on the host cores.
  e2e           the same metric through the C ABI with HOST buffers (H2D of the time grid and
                D2H of all states and error codes inside the timed region), with the raw D2H
                ceiling of the box measured in the same run (d2h_ceiling_gbs, frac_of_ceiling).
  e2e_dropin    the drop-in call itself, perturb::SatelliteBatch::propagate into a
                std::vector<StateVector> (tests/cpp/dropin_bench.cpp, a C++ program).
  workloads     BASELINE.json configs 3, 4, 5 at full size (device-resident), each with its own
                clocks and FP64-pipe share; c5_strong at N > 1 (1 M satellites / N per GPU).
  e2e_rate_balanced  (N > 1) the e2e arm with satellite ranges proportional to each rank's
                measured D2H rate (the GPUs of a box do not share the host links evenly).
  inproc_sharded  (N > 1, rank 0) one process driving all N GPUs: ShardedSatelliteBatch,
                bit-identity against the unsharded batch, device-resident gather over NVLink.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# flops-v0 convention of SURVEY.md section 8d (frozen; change only together with that table)
F_ALG = {0: 1170.0, 1: 1090.0, 2: 1502.0, 3: 1586.0, 4: 2097.0}
BYTES_PER_EVAL = 52.0  # 48 B r,v + 4 B Sgp4Error written; reads amortise to < 1 B
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
N_SM = 148

WORKLOADS = {
    # name: (synthetic config, satellites per GPU, times) -- BASELINE.json configs[1..4];
    # all outputs stay device-resident (c3: 52 GB, c4: 26 GB, c5: 65 GB per GPU of 180 GB)
    "c2": (2, 30000, 1440),
    "c3": (3, 100000, 10000),
    "c4": (4, 50000, 10080),
    "c5": (5, 125000, 10000),
    # time-chunks of the above for quick kernel sweeps
    "c3s": (3, 100000, 2000),
    "c4s": (4, 50000, 2016),
    "c5s": (5, 125000, 2000),
    # small cuts for ncu --set full (ncu saves and restores every output byte per replay pass)
    "c3p": (3, 100000, 512),
    "c4p": (4, 50000, 1024),
}
WORKLOAD_TEXT = {
    "c2": "30k near-Earth synthetic LEO catalogue x 1440 one-minute steps per GPU",
    "c3": "100k mixed near-Earth + deep-space synthetic catalogue x 10k one-minute steps per GPU",
    "c4": "deep-space stress: 25k GEO + 25k Molniya resonant TLEs x 7-day one-minute grid per GPU",
    "c5": "1M-satellite mega-constellation x 10k steps sharded by satellite range, 125k per GPU",
    "c3s": "100k mixed catalogue x 2000 steps per GPU (time-chunk of c3)",
    "c4s": "50k GEO + Molniya catalogue x 2016 steps per GPU (time-chunk of c4)",
    "c5s": "125k mega-constellation satellites x 2000 steps per GPU (time-chunk of c5)",
    "c3p": "100k mixed catalogue x 512 steps (profiling cut of c3)",
    "c4p": "50k GEO + Molniya catalogue x 1024 steps (profiling cut of c4)",
}
E2E_MAX_TIMES = 2000  # host-buffer arm: at most this many times per step (bounds pinned memory)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.power = []
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
            "power_w_max": max(self.power) if self.power else None,
        }


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore the first-touch placement of the pinned host buffers it
    allocates next) to the NUMA node the GPU hangs off: D2H then stays on the local root
    complex / memory controller. Returns a short description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return "numa: node %d has no allowed cpus" % node
        os.sched_setaffinity(0, allowed)
        return "numa node %d, %d cpus" % (node, len(allowed))
    except Exception as exc:  # pragma: no cover - depends on the box
        return "numa: unbound (%s)" % type(exc).__name__


def catalogue_lines(workload, rank, world, total=None):
    """Global catalogue of world x n satellites (or `total`), this rank's satellite range."""
    import perturb_b200 as pb
    cfg, n, m = WORKLOADS[workload]
    total = n * world if total is None else total
    b1, b2 = pb.synth_catalog(cfg, total)
    s1, s2, n_local = pb.shard_blocks(b1, b2, total, rank, world)
    return s1, s2, n_local, m


def config_of(workload, n, m, world=1, cls=None):
    """The `config` object of a bench line: identical keys in the GPU arm and the reference arm."""
    cfg = {
        "workload": WORKLOAD_TEXT[workload],
        "satellites_per_gpu": n, "times": m, "evals_per_step": float(n) * m * world,
        "sharding": "satellite ranges, one process per GPU, no data-path collective",
        "l2": "256 MiB buffer rewritten between timed iterations (outside the event pair); "
              "each step also writes %.2f GB of outputs" % (float(n) * m * 52 / 1e9),
        "outputs": "device-resident SoA planes [6][n][m] f64 + err [n][m] i32",
    }
    if cls is not None:
        cfg["orbit_classes"] = {"near_full": int(cls[0]), "near_simplified": int(cls[1]),
                                "deep_nonresonant": int(cls[2]), "deep_24h": int(cls[3]),
                                "deep_12h": int(cls[4])}
    return cfg


def cpu_arm(workload, steps, warmup, cores=None, budget_s=12.0):
    """The reference's scalar loop on all host cores over a bounded sample of the workload.
    Touches only libperturb_benchtools.so (catalogue text, time grid) and oracle/ -- never the
    product library."""
    import perturb_b200 as pb
    from oracle import port, ref
    cfg, n, m = WORKLOADS[workload]
    b1, b2, n, m = catalogue_lines(workload, 0, 1)
    cores = cores or (os.cpu_count() or 1)
    jd, fr = pb.synth_time_grid(m)
    L1, L2 = pb.block_to_lines(b1, n), pb.block_to_lines(b2, n)
    use_ref = ref.available()

    def make(k):
        if use_ref:
            return ref.RefSatellites.from_tle_lines(L1[:k], L2[:k])
        return port.OracleSatellites.from_tle_lines(L1[:k], L2[:k], 1, "i", flavor=1)

    probe_n = min(n, 64 * cores)
    sats = make(probe_n)
    _, _, _, secs = sats.propagate_grid(jd, fr, use_jd=True, nthreads=cores, want_state=False)
    rate = probe_n * m / max(secs, 1e-9)
    k = int(min(n, max(probe_n, rate * budget_s / max(1, steps + warmup) / m)))
    sats = make(k)
    times = []
    for it in range(warmup + steps):
        _, _, _, secs = sats.propagate_grid(jd, fr, use_jd=True, nthreads=cores, want_state=True)
        if it >= warmup:
            times.append(secs)
    value = k * m * len(times) / sum(times)
    return {
        "value": value,
        "unit": "evals/s",
        "cores": cores,
        "kind": "reference" if use_ref else "port",
        "sample": "%d of %d satellites x all %d times per step, %d steps (walked in time order, "
                  "one contiguous satellite slice per thread)" % (k, n, m, len(times)),
        "ms_per_step": 1e3 * sum(times) / len(times),
        "sample_evals": k * m,
    }


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    base = cpu_arm(args.workload, args.steps, args.warmup, budget_s=60.0)
    cfg, n, m = WORKLOADS[args.workload]
    line = {
        "impl": "reference",
        "metric": "SGP4 evals/sec (sat x time)",
        "value": base["value"],
        "unit": "evals/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args.workload, n, m, max(1, args.gpus)),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def load_profile_counts():
    """FP64 instructions per evaluation of each class kernel as counted by ncu (profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        return {}


def fp64_pipe_share(prof, cls, m, launch_ms, sm_mhz):
    """Share of the FP64 pipe's issue cycles (2 warp instructions per clock per SM) a step keeps
    busy, from the per-class FP64 instruction counts -- the executed-work roofline view."""
    ipe = prof.get("fp64_instr_per_eval")
    if not ipe or not sm_mhz:
        return None
    instr = float(sum(ipe[c] * cls[c] for c in range(5))) * m / 32.0
    cap = launch_ms * 1e-3 * sm_mhz * 1e6 * N_SM * 2.0
    return {"fp64_warp_instr_per_step": instr, "frac_of_fp64_issue_cycles": instr / cap,
            "fp64_instr_per_eval": ipe, "source": prof.get("_fp64_instr_source")}


def time_device_resident(pb, torch, batch, n, m, dev, local, steps, warmup, dist, chunk=None):
    """K timed steps of the hot path with device-resident outputs: CUDA events on the launching
    stream around each step, L2 flushed between steps (outside the event pair), NVML clock
    sampling during the region. A step covers the whole time grid; with `chunk` it is walked in
    time-chunks of that many steps into one reused output buffer (grids whose 52 B/evaluation
    exceed the GPU's memory). Returns per-step ms (this rank), clocks, launches per step."""
    jd, fr = pb.synth_time_grid(m)
    d_jd = torch.tensor(jd, device=dev)
    d_fr = torch.tensor(fr, device=dev)
    mc = m if chunk is None else min(m, chunk)
    out = torch.empty((6, n, mc), dtype=torch.float64, device=dev)
    err = torch.empty((n, mc), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    launches = [0]

    def step():
        launches[0] = 0
        for k0 in range(0, m, mc):
            mk = min(mc, m - k0)
            batch.propagate(d_jd[k0:k0 + mk], d_fr[k0:k0 + mk], out[:, :, :mk], err[:, :mk],
                            stream=sp, sync=False)
            launches[0] += batch.last_launch_count()

    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize(dev)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    # keep the GPU under the same load a little longer if the region was too short to sample
    while len(sampler.samples) < 5 and time.perf_counter() - wall0 < 3.0:
        step()
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    if dist is not None:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    ok_fraction = float((err == 0).double().mean())
    del out, err, flush
    torch.cuda.empty_cache()
    return {"step_ms": step_ms, "clocks": clocks, "launches": launches[0], "wall": wall,
            "ok_fraction": ok_fraction, "chunk": mc}


def extra_workload(pb, torch, name, rank, world, local, dev, dist, fp64_peak, prof, steps,
                   total=None, chunk_bytes=60e9):
    """One more BASELINE config, device-resident, on this rank's shard; returns its record (on
    every rank; values are whole-job, max over ranks)."""
    b1, b2, n, m = catalogue_lines(name, rank, world, total)
    batch = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", device=local, n=n)
    cls = np.bincount(batch.orbit_class(), minlength=5)
    chunk = None
    if float(n) * m * BYTES_PER_EVAL > chunk_bytes:
        chunk = max(128, int(chunk_bytes / (float(n) * BYTES_PER_EVAL)) // 128 * 128)
    r = time_device_resident(pb, torch, batch, n, m, dev, local, steps, 3, dist, chunk)
    batch.close()
    total_ms = float(sum(r["step_ms"]))
    max_ms = pb.max_over_ranks(total_ms, dist, dev)
    n_all = n * world if total is None else total
    value = float(n_all) * m * steps / (max_ms * 1e-3)
    launch_ms = total_ms / steps
    falg = float(sum(F_ALG[c] * cls[c] for c in range(5)) / max(1, cls.sum()))
    tfl = falg * float(n) * m / (launch_ms * 1e-3) * 1e-12
    return {
        "workload": WORKLOAD_TEXT[name], "satellites_per_gpu": n, "satellites": n_all, "times": m,
        "value": value, "unit": "evals/s", "ms_per_step": max_ms / steps, "steps": steps,
        "time_chunk": r["chunk"], "launches_per_step": r["launches"],
        "classes": cls.tolist(), "tflops_v0": tfl,
        "frac_of_fp64_peak_v0": tfl / fp64_peak if fp64_peak > 0 else None,
        "fp64_pipe": fp64_pipe_share(prof, cls, m, launch_ms, r["clocks"].get("sm_mhz")),
        "hbm_gbs": BYTES_PER_EVAL * float(n) * m / (launch_ms * 1e-3) * 1e-9,
        "clocks": r["clocks"], "ok_fraction": r["ok_fraction"],
    }


def nvlink_rx_kib(index):
    """Total NVLink data received by GPU `index` so far (KiB, all links), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        v = pynvml.nvmlDeviceGetFieldValues(
            h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])[0]
        if v.nvmlReturn != 0:
            return None
        return int(v.value.ullVal)
    except Exception:
        return None


def inproc_sharded(pb, torch, workload, world, steps):
    """(rank 0, N > 1) ONE process drives all N GPUs: ShardedSatelliteBatch over the same global
    catalogue, bit-identity against the unsharded batch, and the device-resident gather (every
    shard's kernel 2 stores its rows into GPU 0 over NVLink peer access), with the NVLink
    receive counter of GPU 0 read around it."""
    cfg, n, m = WORKLOADS[workload]
    total = n * world
    b1, b2 = pb.synth_catalog(cfg, total)
    tles, _ = pb.parse_tle_blocks(b1, b2, total, flavor=1)
    jd, fr = pb.synth_time_grid(m)
    dev0 = torch.device("cuda", 0)
    whole = pb.SatelliteBatch.from_parsed(tles, pb.WGS72, "i", device=0)
    ref_out = torch.empty((6, total, m), dtype=torch.float64, device=dev0)
    ref_err = torch.empty((total, m), dtype=torch.int32, device=dev0)
    d_jd, d_fr = torch.tensor(jd, device=dev0), torch.tensor(fr, device=dev0)
    torch.cuda.synchronize(dev0)
    whole.propagate(d_jd, d_fr, ref_out, ref_err)
    whole.close()
    sharded = pb.ShardedSatelliteBatch.from_parsed(tles, list(range(world)), pb.WGS72, "i")
    out, err = sharded.propagate_to_device(jd, fr, device=0)  # warm-up: enables peer access
    same = bool(torch.equal(out.view(torch.int64), ref_out.view(torch.int64)) and
                torch.equal(err, ref_err))
    del ref_out, ref_err
    for d in range(world):
        torch.cuda.synchronize(torch.device("cuda", d))
    rx0 = nvlink_rx_kib(0)
    t0 = time.perf_counter()
    for _ in range(steps):
        sharded.propagate(jd, fr, out, err)  # enqueue on every device, then synchronise all
    secs = (time.perf_counter() - t0) / steps
    rx1 = nvlink_rx_kib(0)
    remote = sum(hi - lo for (lo, hi) in sharded.ranges[1:])
    remote_bytes = float(remote) * m * BYTES_PER_EVAL
    rec = {
        "devices": world, "satellites": total, "times": m, "bit_identical": same,
        "ranges": [list(r) for r in sharded.ranges],
        "evals_per_s_gathered": float(total) * m / secs, "ms_per_step": secs * 1e3, "steps": steps,
        "gather_gbs_into_target": remote_bytes / secs * 1e-9,
        "remote_bytes_per_step_expected": remote_bytes,
        "nvlink_rx_bytes_per_step_gpu0": None if rx0 is None or rx1 is None
        else (rx1 - rx0) * 1024.0 / steps,
        "what": "one process, %d GPUs: every shard's kernel 2 stores its rows straight into "
                "GPU 0 (NVLink peer access); wall clock around enqueue-all + synchronise-all" % world,
    }
    for b in sharded.batches:
        b.close()
    return rec


def dropin_harness(workload, local, steps):
    """perturb::SatelliteBatch::propagate into a std::vector<StateVector>: the C++ harness
    (tests/cpp/dropin_bench.cpp, prebuilt by __graft_entry__.build(); compiled here if absent)."""
    import __graft_entry__ as ge
    exe = ge.DROPIN_BENCH
    if not os.path.exists(exe):
        ge.build_dropin_bench()
    cfg, n, m = WORKLOADS[workload]
    me = min(m, E2E_MAX_TIMES)
    out = subprocess.run([exe, str(cfg), str(n), str(me), str(steps), str(local)],
                         capture_output=True, text=True, timeout=600)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else "{}"
    rec = json.loads(line)
    if out.returncode != 0 or "error" in rec:
        raise RuntimeError("dropin_bench failed (%d): %s %s" % (out.returncode, line, out.stderr[-300:]))
    return rec


def e2e_host_planes(pb, torch, batch, n, m, world, local, dev, dist, steps, bind=True):
    """The hot path through the C ABI with HOST buffers: H2D of the time grid, kernel 2, D2H of
    all states and codes into pinned memory inside the timed region (wall clock, max over ranks).
    Returns the record and the pinned grid / code buffers for the next arm."""
    numa = bind_to_gpu_numa_node(local) if bind else "numa: binding off"
    me = min(m, E2E_MAX_TIMES)
    jd, fr = pb.synth_time_grid(m)
    h_jd = torch.tensor(jd[:me]).pin_memory()
    h_fr = torch.tensor(fr[:me]).pin_memory()
    h_out = torch.empty((6, n, me), dtype=torch.float64).pin_memory()
    h_err = torch.empty((n, me), dtype=torch.int32).pin_memory()
    batch.propagate(h_jd, h_fr, h_out, h_err)  # warm-up: allocates the staging buffers
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    te = time.perf_counter()
    for _ in range(steps):
        batch.propagate(h_jd, h_fr, h_out, h_err)  # synchronises: results are on the host
        _ = int(h_err[0, 0])
    own_s = time.perf_counter() - te
    e2e_s = pb.max_over_ranks(own_s, dist, dev)
    rec = {
        "value": float(n) * me * world * steps / e2e_s,
        "unit": "evals/s",
        "h2d_bytes_per_step": int(2 * me * 8),
        "d2h_bytes_per_step": int(float(n) * me * 52),
        "times_per_step": me,
        "steps": steps,
        "host_buffers": "pinned",
        "placement": numa,
        "d2h_gbs": float(n) * me * world * steps * BYTES_PER_EVAL / e2e_s * 1e-9,
        "d2h_gbs_this_rank": float(n) * me * steps * BYTES_PER_EVAL / own_s * 1e-9,
        "ok_fraction": float((h_err == 0).double().mean()),
    }
    del h_out
    return rec, h_jd, h_fr, h_err


def d2h_ceiling(pb, torch, world, local, dev, dist, window_s=1.0):
    """Raw D2H ceiling of the box: after a barrier every rank copies 256 MiB pieces device ->
    pinned host back to back for the same `window_s` seconds; a rank's rate is the bytes it
    landed in that window, the aggregate is the sum over ranks -- what N GPUs can put into host
    memory when all of them copy all the time. (Summing rates that each rank measured over its
    own copies overstates it: the ranks' copies need not overlap.) Equal satellite shards finish
    with the slowest GPU, so the bound for the e2e arm is N x the slowest rank's rate."""
    piece = 256 << 20
    d = torch.empty(piece, dtype=torch.uint8, device=dev)
    d.fill_(1)
    h = torch.empty(piece, dtype=torch.uint8).pin_memory()
    h.copy_(d, non_blocking=True)  # first touch, untimed
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    done = 0
    while True:
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize(dev)
        now = time.perf_counter() - t0
        if now > window_s:
            break
        done += piece
    own = done / window_s * 1e-9
    slowest = -pb.max_over_ranks(-own, dist, dev)
    total = own
    if dist is not None:
        t = torch.tensor([own], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total = float(t.item())
    del d, h
    return {"_own_gbs": own,
            "d2h_ceiling_gbs": total,
            "d2h_ceiling_gbs_slowest_gpu": slowest,
            "d2h_ceiling_gbs_equal_shards": slowest * world,
            "d2h_ceiling_how": "all %d ranks at once, after a barrier: 256 MiB cudaMemcpyAsync pieces "
                               "device -> pinned host back to back for %.1f s each; bytes landed in the "
                               "common window, summed over ranks; equal_shards = N x the slowest rank "
                               "(equal satellite shards finish with the slowest GPU)" % (world, window_s)}


def e2e_rate_balanced(pb, torch, workload, rank, world, local, dev, dist, steps, own_rate):
    """(N > 1) The pinned-host arm once more over the SAME global catalogue (N x 30 000
    satellites), split into contiguous satellite ranges proportional to the D2H rate each rank
    measured while all ranks were copying (d2h_ceiling): the GPUs of a box do not share the host
    links evenly (e.g. three behind one bridge, one alone), and with equal ranges the whole job
    waits for the slowest. Same calls, same bytes in total; reported BESIDE `e2e` (equal ranges),
    not instead of it."""
    cfg, n, m = WORKLOADS[workload]
    total = n * world
    me = min(m, E2E_MAX_TIMES)
    rates = [None] * world
    dist.all_gather_object(rates, float(own_rate))
    ranges = pb.rate_ranges(total, rates)
    lo, hi = ranges[rank]
    k = hi - lo
    b1, b2 = pb.synth_catalog(cfg, total)
    stride = pb.capi.SYNTH_LINE_STRIDE
    batch = pb.SatelliteBatch.from_text(b1[lo * stride: hi * stride], b2[lo * stride: hi * stride],
                                        pb.WGS72, "i", device=local, n=k)
    jd, fr = pb.synth_time_grid(m)
    h_jd = torch.tensor(jd[:me]).pin_memory()
    h_fr = torch.tensor(fr[:me]).pin_memory()
    h_out = torch.empty((6, k, me), dtype=torch.float64).pin_memory()
    h_err = torch.empty((k, me), dtype=torch.int32).pin_memory()
    batch.propagate(h_jd, h_fr, h_out, h_err)  # warm-up: staging buffers
    dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        batch.propagate(h_jd, h_fr, h_out, h_err)
        _ = int(h_err[0, 0])
    secs = pb.max_over_ranks(time.perf_counter() - t0, dist, dev)
    ok = float((h_err == 0).double().mean())
    batch.close()
    del h_out, h_err
    return {
        "value": float(total) * me * steps / secs, "unit": "evals/s",
        "what": "the e2e arm over the same global catalogue with satellite ranges proportional to "
                "each rank's measured D2H rate (the job no longer waits for the GPUs behind the "
                "busiest host bridge)",
        "satellites_per_rank": [b - a for a, b in ranges],
        "d2h_rate_per_rank_gbs": [float(x) for x in rates],
        "h2d_bytes_per_step": int(2 * me * 8) * world,
        "d2h_bytes_per_step": int(float(total) * me * 52),
        "d2h_gbs": float(total) * me * steps * BYTES_PER_EVAL / secs * 1e-9,
        "times_per_step": me, "steps": steps, "ok_fraction_this_rank": ok,
    }


def add_ceiling_fractions(e2e):
    if e2e.get("d2h_ceiling_gbs_equal_shards"):
        e2e["frac_of_ceiling"] = e2e["d2h_gbs"] / e2e["d2h_ceiling_gbs_equal_shards"]
    if e2e.get("d2h_ceiling_gbs"):
        e2e["frac_of_aggregate_ceiling"] = e2e["d2h_gbs"] / e2e["d2h_ceiling_gbs"]


def run_gpu(args):
    import torch
    import perturb_b200 as pb

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    b1, b2, n, m = catalogue_lines(args.workload, rank, world)
    t0 = time.perf_counter()
    tles, status = pb.parse_tle_blocks(b1, b2, n, flavor=1)
    t_parse = time.perf_counter() - t0
    t0 = time.perf_counter()
    batch = pb.SatelliteBatch.from_parsed(tles, pb.WGS72, "i", device=local)
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    batch_text = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", device=local, n=n)
    t_text = time.perf_counter() - t0
    assert np.array_equal(batch_text.last_error(), batch.last_error())
    batch_text.close()
    cls = np.bincount(batch.orbit_class(), minlength=5)
    falg = float(sum(F_ALG[c] * cls[c] for c in range(5)) / max(1, cls.sum()))

    peak = pb.capi.C.c_double(0.0)
    if pb.capi.tools().perturb_gpu_measure_fp64_peak(local, pb.capi.C.byref(peak)) != 0:
        raise SystemExit("bench.py: perturb_gpu_measure_fp64_peak failed")
    fp64_peak = peak.value

    # ---- device-resident timing: K steps, CUDA events on the launching stream ----
    r = time_device_resident(pb, torch, batch, n, m, dev, local, args.steps, args.warmup, dist)
    step_ms, clocks, launches_per_step, wall = r["step_ms"], r["clocks"], r["launches"], r["wall"]
    total_ms = float(sum(step_ms))
    max_ms = pb.max_over_ranks(total_ms, dist, dev)
    evals_step_rank = float(n) * m
    value = evals_step_rank * world * args.steps / (max_ms * 1e-3)

    if args.quick:
        if rank == 0:
            ms = max_ms / args.steps
            print(json.dumps({"quick": True, "lib": os.path.basename(pb.capi.LIB_PATH),
                              "workload": args.workload, "value": value, "ms_per_step": ms,
                              "tflops_v0": falg * evals_step_rank / (total_ms / args.steps * 1e-3) * 1e-12,
                              "fp64_peak": fp64_peak, "clocks": clocks,
                              "classes": cls.tolist(), "launches": launches_per_step}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- end to end through the C ABI with HOST buffers ----
    all_cpus = os.sched_getaffinity(0)
    me = min(m, E2E_MAX_TIMES)
    jd, fr = pb.synth_time_grid(m)
    e2e_steps = max(2, min(args.steps, 5))
    e2e, h_jd, h_fr, h_err = e2e_host_planes(pb, torch, batch, n, m, world, local, dev, dist,
                                             e2e_steps, bind=os.environ.get("PERTURB_BENCH_NUMA", "on") != "off")
    codes_ok = e2e.pop("ok_fraction")
    if args.e2e_only:
        e2e.update(d2h_ceiling(pb, torch, world, local, dev, dist))
        own_rate = e2e.pop("_own_gbs")
        add_ceiling_fractions(e2e)
        if world > 1:
            e2e["rate_balanced"] = e2e_rate_balanced(pb, torch, args.workload, rank, world, local,
                                                     dev, dist, e2e_steps, own_rate)
        per_rank = [None] * world
        if dist is not None:
            dist.all_gather_object(per_rank, e2e["d2h_gbs_this_rank"])
        else:
            per_rank = [e2e["d2h_gbs_this_rank"]]
        if rank == 0:
            print(json.dumps({"e2e_only": True, "n_gpus": world, "workload": args.workload,
                              "host_chunks": os.environ.get("PERTURB_GPU_HOST_CHUNKS"),
                              "value_device_resident": value, "e2e": e2e,
                              "d2h_gbs_per_rank": per_rank}))
        batch.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- the same through perturb_gpu_propagate_states: the caller's StateVector[n * m] ----
    h_states = torch.empty((n, me, 8), dtype=torch.float64).pin_memory()
    batch.propagate_states(h_jd, h_fr, h_states, h_err)  # warm-up (staging buffers)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    tr = time.perf_counter()
    for _ in range(e2e_steps):
        batch.propagate_states(h_jd, h_fr, h_states, h_err)  # synchronises
        _ = int(h_err[0, 0])
    rec_s = pb.max_over_ranks(time.perf_counter() - tr, dist, dev)
    rec_value = float(n) * me * world * e2e_steps / rec_s
    del h_states, h_err
    ceiling = d2h_ceiling(pb, torch, world, local, dev, dist)
    own_rate = ceiling.pop("_own_gbs")
    balanced = None
    if world > 1 and not args.no_extra:
        balanced = e2e_rate_balanced(pb, torch, args.workload, rank, world, local, dev, dist,
                                     e2e_steps, own_rate)
    # ---- the fused consumer (N3): host time grid in, one 40-byte summary per satellite out ----
    hs_jd = torch.tensor(jd).pin_memory()
    hs_fr = torch.tensor(fr).pin_memory()
    summ = batch.propagate_summary(hs_jd, hs_fr)  # warm-up: allocates the partials scratch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ts = time.perf_counter()
    for _ in range(e2e_steps):
        summ = batch.propagate_summary(hs_jd, hs_fr)  # synchronises: summaries are on the host
    sum_s = time.perf_counter() - ts
    sum_value = float(n) * m * world * e2e_steps / pb.max_over_ranks(sum_s, dist, dev)
    sum_launches = batch.last_launch_count()
    # ---- two more fused consumers: ground-station visibility and closest approach ----
    jd_np, fr_np = np.asarray(jd), np.asarray(fr)
    stations = [(np.radians(40.0), np.radians(-105.0), 1.6, np.radians(10.0)),
                (0.0, np.radians(30.0), 0.0, np.radians(5.0)),
                (np.radians(78.2), np.radians(15.4), 0.4, np.radians(5.0)),
                (np.radians(-33.9), np.radians(151.2), 0.05, np.radians(10.0))]
    vis, _ = batch.propagate_visibility(jd_np, fr_np, stations)  # warm-up: scratch
    if dist is not None:
        dist.barrier()
    tv = time.perf_counter()
    for _ in range(e2e_steps):
        vis, _ = batch.propagate_visibility(jd_np, fr_np, stations)  # host in, host out
    vis_s = pb.max_over_ranks(time.perf_counter() - tv, dist, dev)
    prim_idx = [0, n // 3, (2 * n) // 3, n - 1]
    d_prim = torch.empty((6, len(prim_idx), m), dtype=torch.float64, device=dev)
    d_perr = torch.empty((len(prim_idx), m), dtype=torch.int32, device=dev)
    sub = (pb.capi.Tle * len(prim_idx))(*[tles[i] for i in prim_idx])
    prim_batch = pb.SatelliteBatch.from_parsed(sub, pb.WGS72, "i", device=local)
    prim_batch.propagate(torch.tensor(jd_np, device=dev), torch.tensor(fr_np, device=dev), d_prim, d_perr)
    prim_batch.close()
    primaries = np.ascontiguousarray(d_prim[0:3].permute(1, 0, 2).cpu().numpy())
    prox = batch.propagate_proximity(jd_np, fr_np, primaries, 200.0)  # warm-up
    if dist is not None:
        dist.barrier()
    tp = time.perf_counter()
    for _ in range(e2e_steps):
        prox = batch.propagate_proximity(jd_np, fr_np, primaries, 200.0)
    prox_s = pb.max_over_ranks(time.perf_counter() - tp, dist, dev)
    screens = {
        "e2e_visibility": {
            "value": float(n) * m * world * e2e_steps / vis_s, "unit": "evals/s",
            "what": "perturb_gpu_propagate_visibility: host time grid in, per (satellite, station) "
                    "pass summaries out, %d stations, elevation masks 5-10 deg" % len(stations),
            "h2d_bytes_per_step": int(2 * m * 8 + 32 * len(stations)),
            "d2h_bytes_per_step": int(n * len(stations) * 32), "steps": e2e_steps,
            "visible_fraction": float(vis["n_visible"].sum()) / (float(n) * m * len(stations)),
            "passes": int(vis["n_passes"].sum()),
        },
        "e2e_proximity": {
            "value": float(n) * m * world * e2e_steps / prox_s, "unit": "evals/s",
            "what": "perturb_gpu_propagate_proximity: host time grid and %d primaries' positions in, "
                    "per (satellite, primary) closest approach out, threshold 200 km" % len(prim_idx),
            "h2d_bytes_per_step": int(2 * m * 8 + primaries.nbytes),
            "d2h_bytes_per_step": int(n * len(prim_idx) * 16), "steps": e2e_steps,
            "pairs_below_threshold": int((prox["n_below"] > 0).sum()),
        },
    }
    batch.close()
    torch.cuda.empty_cache()
    # ---- the drop-in call itself: C++ user code, std::vector destinations (one process per GPU)
    dropin = None
    try:
        if dist is not None:
            dist.barrier()
        dropin = dropin_harness(args.workload, local, max(2, min(args.steps, 3)))
        drop_s = pb.max_over_ranks(dropin["dropin_ms"], dist, dev) * 1e-3
        pin_s = pb.max_over_ranks(dropin["pinned_ms"], dist, dev) * 1e-3
        reg_s = pb.max_over_ranks(dropin.get("registered_ms", 0.0), dist, dev) * 1e-3
        dn = float(dropin["satellites"]) * dropin["times"] * world
        dropin = {
            "value": dn / drop_s, "unit": "evals/s",
            "what": "perturb::SatelliteBatch::propagate(times, m, svs, errs) from a C++ program "
                    "(tests/cpp/dropin_bench.cpp): std::vector<StateVector> + std::vector<Sgp4Error> "
                    "destinations (pageable), untouched-on-error semantics, wall clock per call",
            "ms_per_call": drop_s * 1e3,
            "h2d_bytes_per_step": int(2 * dropin["times"] * 8),
            "d2h_bytes_per_step": int(float(dropin["satellites"]) * dropin["times"] * 68),
            "steps": dropin["steps"],
            "pinned_states_value": dn / pin_s if pin_s > 0 else None,
            "ratio_to_pinned_states": pin_s / drop_s if drop_s > 0 else None,
            # the same call after `perturb::PinnedRegion pin(svs.data(), bytes)`: the vectors
            # page-locked in place, the copy engine writes them directly (no bounce copy)
            "registered_value": dn / reg_s if reg_s > 0 else None,
            "registered_ratio_to_pinned_states": pin_s / reg_s if reg_s > 0 else None,
            "register_once_ms": dropin.get("register_once_ms"),
            "registered_equals_dropin_bit_for_bit": dropin.get("registered_equals_dropin"),
            "equals_pinned_states_bit_for_bit": dropin["dropin_equals_pinned"],
            "ok_fraction": dropin["ok_fraction"],
        }
    except Exception as exc:  # the harness is evidence, not the metric: report, do not die
        dropin = {"value": None, "error": "%s: %s" % (type(exc).__name__, exc)}
    os.sched_setaffinity(0, all_cpus)  # the CPU arm below uses every host core again

    # ---- the other BASELINE configs, device-resident, full size ----
    prof = load_profile_counts()
    workloads = {}
    if not args.no_extra:
        xsteps = 3
        if world == 1:
            for name in ("c3", "c4", "c5"):
                workloads[name] = extra_workload(pb, torch, name, rank, world, local, dev, dist,
                                                 fp64_peak, prof, xsteps)
        else:
            # strong scaling of config 5 as BASELINE.json states it: 1 M satellites x 10 000
            # times over N GPUs, time-chunked where 52 B x evaluations exceed the device memory
            workloads["c5_strong"] = extra_workload(pb, torch, "c5", rank, world, local, dev, dist,
                                                    fp64_peak, prof, xsteps, total=1000000)
            workloads["c5_strong"]["scaling"] = "strong"
    inproc = None
    if world > 1 and not args.no_extra:
        if rank == 0:
            try:
                inproc = inproc_sharded(pb, torch, args.workload, world, 5)
            except Exception as exc:
                inproc = {"error": "%s: %s" % (type(exc).__name__, exc)}
        dist.barrier()  # the other ranks wait here, their GPUs idle

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms_per_step = max_ms / args.steps
    launch_ms = total_ms / args.steps  # rank 0's own average launch group
    achieved_tflops = falg * evals_step_rank / (launch_ms * 1e-3) * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = prof.get(args.workload)
    pipe = fp64_pipe_share(prof, cls, m, launch_ms, clocks.get("sm_mhz"))
    cpu = cpu_arm(args.workload, 2, 1)
    e2e.update(ceiling)
    add_ceiling_fractions(e2e)
    line = {
        "metric": "SGP4 evals/sec (sat x time)",
        "value": value,
        "unit": "evals/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args.workload, n, m, world),
        "orbit_classes": {"near_full": int(cls[0]), "near_simplified": int(cls[1]),
                          "deep_nonresonant": int(cls[2]), "deep_24h": int(cls[3]),
                          "deep_12h": int(cls[4])},
        "roofline": {
            "bound": "fp64",
            "achieved": achieved_tflops,
            "peak": fp64_peak,
            "unit": "TFLOP/s",
            "frac": achieved_tflops / fp64_peak if fp64_peak > 0 else None,
            "traffic": traffic,
            "peak_source": "DFMA microbenchmark measured in this run (perturb_gpu_measure_fp64_peak)",
            "peak_nominal": FP64_NOMINAL_TFLOPS,
            "frac_of_nominal": achieved_tflops / FP64_NOMINAL_TFLOPS,
            "flops_per_eval": falg,
            "kernel": "sgp4_prop_kernel (all class launches of one step)",
            "launch_ms": launch_ms,
            "note": "flops-v0 charges every sin/cos/div/sqrt at its libdevice fast-path cost "
                    "(SURVEY.md 8d); the kernel needs fewer FP64 operations for the same "
                    "evaluation, so frac can exceed 1 -- fp64_pipe is the executed-work view",
            "fp64_pipe": pipe,
        },
        "roofline_hbm": {
            "bound": "hbm",
            "achieved": BYTES_PER_EVAL * evals_step_rank / (launch_ms * 1e-3) * 1e-9,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": BYTES_PER_EVAL * evals_step_rank / (launch_ms * 1e-3) * 1e-9 / hbm_peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
        },
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": e2e,
        "e2e_states": {
            "value": rec_value,
            "unit": "evals/s",
            "what": "perturb_gpu_propagate_states: host time grid in, the caller's array of "
                    "64-byte perturb::StateVector records + error codes out (pinned host)",
            "h2d_bytes_per_step": int(2 * me * 8),
            "d2h_bytes_per_step": int(float(n) * me * 68),
            "steps": e2e_steps,
        },
        "e2e_dropin": dropin,
        "e2e_rate_balanced": balanced,
        "e2e_summary": {
            "value": sum_value,
            "unit": "evals/s",
            "what": "perturb_gpu_propagate_summary, host time grid in / host per-satellite "
                    "summaries out, full grid; wall clock incl. copies and synchronisation",
            "h2d_bytes_per_step": int(2 * m * 8),
            "d2h_bytes_per_step": int(n * 40),
            "steps": e2e_steps,
            "launches_per_step": int(sum_launches),
            "ok_fraction": float(summ["n_ok"].sum()) / (float(n) * m),
        },
        "e2e_visibility": screens["e2e_visibility"],
        "e2e_proximity": screens["e2e_proximity"],
        "workloads": workloads,
        "inproc_sharded": inproc,
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "init": {"satellites": n, "host_parse_ms": 1e3 * t_parse,
                 "sgp4init_and_sort_ms": 1e3 * t_init,
                 "device_parse_sgp4init_sort_ms": 1e3 * t_text},
        "wall_s_timed_region": wall,
        "ok_fraction": codes_ok,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--quick", action="store_true",
                    help="device-resident timing only (no e2e, no CPU arm): kernel tuning sweeps")
    ap.add_argument("--e2e-only", action="store_true",
                    help="device-resident timing + the pinned-host arm + the D2H ceiling only "
                         "(multi-GPU host-ingest diagnostics; PERTURB_BENCH_NUMA=off skips the binding)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the other BASELINE configs and the in-process multi-GPU check")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
