"""Ad-hoc first-light check on a GPU box: verification catalogue + small synthetic sets
against the CPU oracle port. (The real parity suite is tests/test_gpu_parity.py.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import perturb_b200 as pb
from oracle import port

VER = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "SGP4-VER.TLE")


def read_ver(path):
    l1, l2 = [], []
    lines = [ln.rstrip("\n") for ln in open(path)]
    it = iter(lines)
    for a in it:
        if a.startswith("#"):
            continue
        l1.append(a); l2.append(next(it))
    return l1, l2


def compare(tag, b, O, mins=None, jd=None, fr=None):
    t0 = time.time()
    if mins is not None:
        pv, err = b.propagate_from_epoch(mins)
        eo, ro, vo, _ = O.propagate_grid(mins, use_jd=False, nthreads=8)
    else:
        pv, err = b.propagate(jd, fr)
        eo, ro, vo, _ = O.propagate_grid(jd, fr, use_jd=True, nthreads=8)
    pos = np.moveaxis(pv[0:3], 0, -1); vel = np.moveaxis(pv[3:6], 0, -1)
    code_bad = (err != eo)
    ok = (eo == 0) | (eo == 6)
    dr = np.abs(pos - ro).max(axis=-1); dv = np.abs(vel - vo).max(axis=-1)
    nanmis = (np.isnan(pos).any(-1) != np.isnan(ro).any(-1)).sum()
    m = ok & ~code_bad
    print("%s: evals %d  code mismatches %d  nan-pattern mismatches %d  max|dr| %.3e km  max|dv| %.3e km/s  codes %s  (%.1fs)" % (
        tag, err.size, code_bad.sum(), nanmis, np.nanmax(np.where(m, dr, 0)), np.nanmax(np.where(m, dv, 0)),
        np.bincount(eo.ravel(), minlength=7).tolist(), time.time() - t0))
    if code_bad.any():
        idx = np.argwhere(code_bad)[:10]
        for i, k in idx:
            print("   code mismatch sat %d time %d gpu %d cpu %d" % (i, k, err[i, k], eo[i, k]))
    if m.any():
        i, k = np.unravel_index(np.nanargmax(np.where(m, dr, 0)), dr.shape)
        print("   worst dr at sat %d time %d: gpu %s cpu %s" % (i, k, pos[i, k], ro[i, k]))
    return code_bad.sum(), np.nanmax(np.where(m, dr, 0)), np.nanmax(np.where(m, dv, 0))


def main():
    l1, l2 = read_ver(VER)
    l1s = [a[:69] for a in l1]; l2s = [a[:69] for a in l2]
    grids = [tuple(float(x) for x in b[69:].split()) for b in l2]
    times = set([0.0, 0.5, 5.0, 30.0, 1440.0, 20000.0])
    for (s, e, d) in grids:
        t = s
        times.add(t)
        while t < e:
            t = min(t + d, e); times.add(t)
    mins = np.array(sorted(times))
    for ops in "ai":
        b = pb.SatelliteBatch.from_tles(l1s, l2s, pb.WGS72, ops)
        O = port.OracleSatellites.from_tle_lines(l1s, l2s, 1, ops, flavor=1)
        print("opsmode", ops, "init errors equal:", np.array_equal(b.last_error(), O.init_error), b.last_error().tolist())
        print(" classes", b.orbit_class().tolist())
        # record parity
        names = b.record_field_names()
        worst = 0
        for nm in names:
            g = b.record_field(nm)
            if nm == "method_d": o = O.field("deep").astype(float)
            else: o = O.field(nm).astype(float)
            den = np.maximum(np.abs(o), 1e-300)
            rel = np.where(o == g, 0, np.abs(g - o) / den)
            if rel.max() > 1e-9:
                print("   field", nm, "max rel", rel.max(), "at", rel.argmax(), g[rel.argmax()], o[rel.argmax()])
            worst = max(worst, rel.max())
        print(" worst record rel diff", worst)
        compare("VER mins ops=" + ops, b, O, mins=mins)
    for cfg, n, m in ((2, 3000, 480), (3, 4000, 700), (4, 2000, 1000), (5, 3000, 300)):
        b1, b2 = pb.synth_catalog(cfg, n)
        L1 = pb.block_to_lines(b1, n); L2 = pb.block_to_lines(b2, n)
        b = pb.SatelliteBatch.from_tles(L1, L2)
        O = port.OracleSatellites.from_tle_lines(L1, L2, 1, "i", flavor=1)
        print("cfg", cfg, "init errors equal:", np.array_equal(b.last_error(), O.init_error), "classes", np.bincount(b.orbit_class(), minlength=5).tolist())
        jd, fr = pb.synth_time_grid(10080)
        sel = np.linspace(0, 10079, m).astype(int)
        compare("C%d jd" % cfg, b, O, jd=jd[sel], fr=fr[sel])


if __name__ == "__main__":
    main()
