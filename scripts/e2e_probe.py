"""Alternates the two host-buffer paths (structure-of-arrays planes / StateVector records) of
config 2 to separate code effects from run-to-run PCIe variation. Prints GB/s per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import perturb_b200 as pb

n, m = 30000, 1440
b1, b2 = pb.synth_catalog(2, n)
batch = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", device=0, n=n)
jd, fr = pb.synth_time_grid(m)
h_jd = torch.tensor(jd).pin_memory(); h_fr = torch.tensor(fr).pin_memory()
h_out = torch.empty((6, n, m), dtype=torch.float64).pin_memory()
h_err = torch.empty((n, m), dtype=torch.int32).pin_memory()
h_st = torch.empty((n, m, 8), dtype=torch.float64).pin_memory()
d_out = torch.empty((6, n, m), dtype=torch.float64, device="cuda")
for rep in range(4):
    for name, fn, nbytes in (("planes", lambda: batch.propagate(h_jd, h_fr, h_out, h_err), n * m * 52),
                             ("records", lambda: batch.propagate_states(h_jd, h_fr, h_st, h_err), n * m * 68)):
        fn()
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        print("%-8s %.1f ms  %.1f GB/s" % (name, dt * 1e3, nbytes / dt * 1e-9))
    # raw copies of the same sizes for reference
    for name, dst, src in (("raw 2.07GB", h_out, d_out),):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%-10s %.1f ms  %.1f GB/s" % (name, dt * 1e3, dst.numel() * 8 / dt * 1e-9))
