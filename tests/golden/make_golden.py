"""Generates the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref/libperturb_ref.so, built from /root/reference by oracle/Makefile).

Run in the build container only (the reference tree does not exist on the GPU box):

    make -C oracle && python tests/golden/make_golden.py

Outputs (committed):
    sgp4_ver_golden.npz   SGP4-VER.TLE through twoline2rv('v', opsmode) exactly as
                          tests/test_perturb.cpp:28-50,343-398 drives it, both opsmodes:
                          init errors, the full elsetrec numerics, err/r/v on every satellite's
                          own verification grid plus the six extra times of :533, and the
                          stop-at-first-error walk (the 640 evaluations of SURVEY.md 8c);
                          the ISS example of examples/cmake-local/main.cpp:11-26; both TLE
                          parsers' numeric output for every verification TLE.
    synth_golden.npz      the first satellites of each synthetic catalogue (configs 2-5)
                          through Satellite::from_tle + Satellite::propagate(JulianDate).
SGP4-VER.TLE itself is a copy of the reference's input fixture (data, not source).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
import perturb_b200 as pb  # noqa: E402  (synthetic catalogue text only; no GPU needed)

EXTRA_TIMES = [0.0, 0.5, 5.0, 30.0, 1440.0, 20000.0]  # tests/test_perturb.cpp:533


def verification_grids(l2):
    return [tuple(float(x) for x in b[69:].split()) for b in l2]


def union_grid(grids):
    times = set(EXTRA_TIMES)
    for (s, e, d) in grids:
        t = s
        times.add(t)
        while t < e:
            t = min(t + d, e)
            times.add(t)
    return np.array(sorted(times))


def full_grid(sats, grids):
    """Every time of every satellite's own grid plus EXTRA_TIMES, errors or not.
    Rows: sat, tsince, err, r[3], v[3]."""
    rows = []
    for i, g in enumerate(grids):
        for t in union_grid([g]):
            e, r, v = sats.propagate_mins_one(i, t)
            rows.append((i, t, e, *r, *v))
    return np.array(rows, dtype=np.float64)


def walk(sats, grids):
    """test_sgp4_verification_mode's loop, tests/test_perturb.cpp:352-398."""
    rows = []
    for i, (start, stop, delta) in enumerate(grids):
        e, r, v = sats.propagate_mins_one(i, 0.0)
        rows.append((i, 0.0, e, *r, *v))
        tsince = start
        if abs(tsince) > 1e-8:
            tsince -= delta
        last = e
        while tsince < stop and last == 0:
            tsince = min(tsince + delta, stop)
            last, r, v = sats.propagate_mins_one(i, tsince)
            rows.append((i, tsince, last, *r, *v))
    return np.array(rows, dtype=np.float64)


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle)"
    l1, l2 = ref.read_verif_file(os.path.join(HERE, "SGP4-VER.TLE"))
    l1s = [a[:69] for a in l1]
    l2s = [b[:69] for b in l2]
    out = {"field_names": np.array(ref.FIELD_NAMES)}
    out["lines1"] = np.array(l1)
    out["lines2"] = np.array(l2)
    for ops in "ai":
        sats = ref.RefSatellites.from_verif(l1, l2, ops)
        grids = sats.grids
        assert grids == verification_grids(l2)
        out["init_error_" + ops] = sats.init_error
        out["fields_" + ops] = sats.fields()
        out["full_" + ops] = full_grid(sats, grids)
        fresh = ref.RefSatellites.from_verif(l1, l2, ops)
        out["walk_" + ops] = walk(fresh, grids)
    out["grids"] = np.array(grids)
    # ClassicalOrbitalElements of every written state of the full grids (opsmode 'a')
    rows = out["full_a"]
    ok = np.isin(rows[:, 2].astype(int), (0, 6))
    out["coe_a"] = ref.rv2coe(rows[ok, 3:6], rows[ok, 6:9])

    # public-API route, opsmode 'i': Satellite::from_tle, then propagate(JulianDate)
    pub = ref.RefSatellites.from_tle_lines(l1s, l2s)
    out["pub_init_error"] = pub.init_error
    jd0, fr0 = pub.epoch()
    out["pub_epoch_jd"], out["pub_epoch_frac"] = jd0, fr0
    out["pub_fields"] = pub.fields()

    # both parsers' numbers
    parsed, perr = ref.RefSatellites.from_parsed(l1s, l2s)
    out["tle_parse_error"] = perr
    out["tle_parse_values"] = np.array(
        [[getattr(t, k) for k in pb.capi.TLE_FIELDS] for t in parsed.tles])

    # ISS example, examples/cmake-local/main.cpp:11-26
    iss1 = "1 25544U 98067A   22071.78032407  .00021395  00000-0  39008-3 0  9996"
    iss2 = "2 25544  51.6424  94.0370 0004047 256.5103  89.8846 15.49386383330227"
    iss = ref.RefSatellites.from_tle_lines([iss1], [iss2])
    jd = 2459652.5  # JulianDate(DateTime{2022, 3, 14, 1, 59, 26.535}) via jday_SGP4
    fr = (26.535 + 59 * 60.0 + 1 * 3600.0) / 86400.0
    e, r, v = iss.propagate_jd_one(0, jd, fr)
    out["iss_lines"] = np.array([iss1, iss2])
    out["iss_jd"] = np.array([jd, fr])
    out["iss_err"] = np.array([e])
    out["iss_r"], out["iss_v"] = r, v
    ej, ef = iss.epoch()
    out["iss_epoch"] = np.array([ej[0], ef[0]])
    np.savez_compressed(os.path.join(HERE, "sgp4_ver_golden.npz"), **out)
    print("sgp4_ver_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})

    # synthetic catalogues
    syn = {}
    jd, fr = pb.synth_time_grid(10080)
    sel = np.unique(np.concatenate([np.arange(0, 8), np.linspace(0, 10079, 24).astype(int)]))
    syn["time_index"] = sel
    for cfg, n in ((2, 128), (3, 256), (4, 128), (5, 64)):
        b1, b2 = pb.synth_catalog(cfg, n)
        L1 = pb.block_to_lines(b1, n)
        L2 = pb.block_to_lines(b2, n)
        sats = ref.RefSatellites.from_tle_lines(L1, L2)
        err, pos, vel, _ = sats.propagate_grid(jd[sel], fr[sel], use_jd=True)
        syn["lines1_c%d" % cfg] = np.array(L1)
        syn["lines2_c%d" % cfg] = np.array(L2)
        syn["init_error_c%d" % cfg] = sats.init_error
        syn["err_c%d" % cfg] = err
        syn["pos_c%d" % cfg] = pos
        syn["vel_c%d" % cfg] = vel
        print("cfg", cfg, "codes", np.bincount(err.ravel(), minlength=7).tolist())
    np.savez_compressed(os.path.join(HERE, "synth_golden.npz"), **syn)


if __name__ == "__main__":
    main()
