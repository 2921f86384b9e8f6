"""GPU tests of the fused consumers of SURVEY.md 8f N3 beyond the radius envelope: the
ground-station visibility screen and the closest-approach screen. Each is checked against the
same reduction done in numpy over the ORACLE's full states (the consumer itself writes nothing
of size [n, m]). Integer fields must be exact wherever no evaluation sits within the parity bar
of a decision threshold; such borderline pairs are counted and bounded."""
import numpy as np
import pytest

import perturb_b200 as pb
from helpers import posvel_to_rv
from oracle import port

pytestmark = pytest.mark.gpu

TWOPI = 2.0 * np.pi


def gstime(jdut1):
    """gstime_SGP4, src/sgp4.cpp:2506-2525, in numpy (same operations, C fmod)."""
    tut1 = (jdut1 - 2451545.0) / 36525.0
    temp = (-6.2e-6 * tut1 * tut1 * tut1 + 0.093104 * tut1 * tut1
            + (876600.0 * 3600 + 8640184.812866) * tut1 + 67310.54841)
    temp = np.fmod(temp * (np.pi / 180.0) / 240.0, TWOPI)
    return np.where(temp < 0.0, temp + TWOPI, temp)


def station_frame(lat, lon, alt, a=6378.135, flat=1.0 / 298.26):
    e2 = flat * (2.0 - flat)
    sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
    nn = a / np.sqrt(1.0 - e2 * sl * sl)
    pos = np.array([(nn + alt) * cl * co, (nn + alt) * cl * so, (nn * (1.0 - e2) + alt) * sl])
    up = np.array([cl * co, cl * so, sl])
    return pos, up


def runs_of(bits):
    """(rise, set) index pairs of the maximal runs of True."""
    d = np.diff(np.concatenate([[0], bits.astype(np.int8), [0]]))
    return list(zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1) - 1))


def _setup(cfg, n, m):
    b1, b2 = pb.synth_catalog(cfg, n)
    L1, L2 = pb.block_to_lines(b1, n), pb.block_to_lines(b2, n)
    batch = pb.SatelliteBatch.from_tles(L1, L2)
    orc = port.OracleSatellites.from_tle_lines(L1, L2, 1, "i", flavor=1)
    jd, fr = pb.synth_time_grid(10080)
    sel = np.linspace(0, 10079, m).astype(int)
    return batch, orc, jd[sel], fr[sel]


@pytest.mark.parametrize("cfg,n,m", [(2, 600, 1000), (3, 500, 333), (3, 200, 40)])
def test_visibility_screen_matches_a_reduction_of_the_oracle_states(cfg, n, m):
    batch, orc, jd, fr = _setup(cfg, n, m)
    stations = [(np.radians(40.0), np.radians(-105.0), 1.6, np.radians(10.0)),
                (0.0, np.radians(30.0), 0.0, 0.0),
                (np.radians(78.2), np.radians(15.4), 0.4, np.radians(5.0)),
                (np.radians(-33.9), np.radians(151.2), 0.05, np.radians(25.0))]
    max_passes = 6
    got, passes = batch.propagate_visibility(jd, fr, stations, max_passes=max_passes)
    assert got.shape == (n, len(stations)) and passes.shape == (n, len(stations), max_passes, 2)

    eo, ro, vo, _ = orc.propagate_grid(jd, fr, use_jd=True, nthreads=8)
    theta = gstime(jd + fr)
    s, c = np.sin(theta), np.cos(theta)
    recef = np.stack([c * ro[..., 0] + s * ro[..., 1], c * ro[..., 1] - s * ro[..., 0], ro[..., 2]], -1)
    margin = 1e-8  # 1e-6 km of position over >= 160 km of range, with room
    borderline = 0
    for j, (lat, lon, alt, mask) in enumerate(stations):
        pos, up = station_frame(lat, lon, alt)
        rho = recef - pos
        sin_el = (rho @ up) / np.linalg.norm(rho, axis=-1)
        ok = eo == 0
        vis = ok & (sin_el >= np.sin(mask))
        near = ok & (np.abs(sin_el - np.sin(mask)) < margin)
        for i in range(n):
            g = got[i, j]
            if ok[i].any():
                best = np.where(ok[i], sin_el[i], -np.inf)
                assert abs(g["max_sin_elevation"] - best.max()) < margin
                assert best[g["k_max_elevation"]] > best.max() - margin
            else:
                assert g["k_max_elevation"] == -1 and g["max_sin_elevation"] == -np.inf
            if near[i].any():
                borderline += 1
                continue
            want = runs_of(vis[i])
            assert g["n_visible"] == vis[i].sum(), (i, j)
            assert g["n_passes"] == len(want), (i, j)
            assert g["k_first_rise"] == (want[0][0] if want else -1)
            assert g["k_last_visible"] == (want[-1][1] if want else -1)
            rec = [tuple(x) for x in passes[i, j] if x[0] >= 0]
            assert rec == [(int(a), int(b)) for a, b in want[:max_passes]], (i, j)
            assert (passes[i, j, len(rec):] == -1).all()
    assert got["n_visible"].sum() > 0 and (got["n_passes"] > 1).any()
    assert borderline <= 0.01 * n * len(stations) + 2, borderline
    # the consumer agrees with the GPU's own full output bit-tightly (same evaluation code)
    pv, err = batch.propagate(jd, fr)
    gpos, _ = posvel_to_rv(pv)
    ge = np.stack([c * gpos[..., 0] + s * gpos[..., 1], c * gpos[..., 1] - s * gpos[..., 0],
                   gpos[..., 2]], -1)
    pos, up = station_frame(*stations[0][:3])
    rho = ge - pos
    sin_el = (rho @ up) / np.linalg.norm(rho, axis=-1)
    vis = (err == 0) & (sin_el >= np.sin(stations[0][3]))
    clear = ~(((err == 0) & (np.abs(sin_el - np.sin(stations[0][3])) < 1e-12)).any(1))
    assert np.array_equal(got["n_visible"][clear, 0], vis.sum(1)[clear])


def test_visibility_screen_device_outputs_and_errors():
    import torch
    batch, orc, jd, fr = _setup(3, 300, 200)
    stations = [(0.5, 1.0, 0.0, 0.1)]
    host, _ = batch.propagate_visibility(jd, fr, stations)
    dev = torch.device("cuda", 0)
    out = torch.zeros(300 * batch.PASS_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    st_arr = np.array(stations, dtype=np.float64)
    from perturb_b200 import capi
    status = capi.lib().perturb_gpu_propagate_visibility(
        batch._h, jd.ctypes.data, fr.ctypes.data, jd.size, st_arr.ctypes.data, 1, out.data_ptr(),
        None, 0, None)
    capi.check(status, "perturb_gpu_propagate_visibility")
    batch.synchronize()
    same = out.cpu().numpy().view(batch.PASS_DTYPE)
    assert np.array_equal(same.view(np.uint8), host.reshape(-1).view(np.uint8))
    # the minutes form has no absolute time, hence no sidereal angle: rejected, loudly
    status = capi.lib().perturb_gpu_propagate_visibility(
        batch._h, jd.ctypes.data, None, jd.size, st_arr.ctypes.data, 1, out.data_ptr(), None, 0, None)
    assert status == -1 and b"visibility" in capi.lib().perturb_gpu_strerror()


@pytest.mark.parametrize("cfg,n,m,thr", [(5, 800, 700, 500.0), (3, 400, 90, 3000.0)])
def test_proximity_screen_matches_a_reduction_of_the_oracle_states(cfg, n, m, thr):
    batch, orc, jd, fr = _setup(cfg, n, m)
    pv, err = batch.propagate(jd, fr)
    prim_idx = [0, 7, n // 2, n - 1]
    primaries = np.ascontiguousarray(pv[0:3, prim_idx, :].transpose(1, 0, 2))  # [P, 3, m]
    got = batch.propagate_proximity(jd, fr, primaries, thr)
    assert got.shape == (n, len(prim_idx))

    eo, ro, vo, _ = orc.propagate_grid(jd, fr, use_jd=True, nthreads=8)
    assert np.array_equal(eo, err)
    margin = 4e-6  # two positions, each within 1e-6 km of the reference, with room
    borderline = 0
    for j, p in enumerate(prim_idx):
        d = np.linalg.norm(ro - np.moveaxis(primaries[j], 0, -1)[None], axis=-1)  # [n, m]
        have = (eo == 0) & np.isfinite(d)
        for i in range(n):
            g = got[i, j]
            if not have[i].any():
                assert g["k_min"] == -1 and np.isinf(g["d_min_km"])
                continue
            dm = np.where(have[i], d[i], np.inf)
            assert abs(g["d_min_km"] - dm.min()) < margin, (i, j)
            assert dm[g["k_min"]] < dm.min() + margin
            near = have[i] & (np.abs(d[i] - thr) < margin)
            want = int((have[i] & (d[i] < thr)).sum())
            if near.any():
                borderline += 1
                assert abs(g["n_below"] - want) <= near.sum()
            else:
                assert g["n_below"] == want, (i, j)
        # a primary is at distance exactly zero from itself at every time it has a state
        assert got[p, j]["d_min_km"] == 0.0 and got[p, j]["n_below"] == (err[p] == 0).sum()
        assert got[p, j]["k_min"] == int(np.flatnonzero(err[p] == 0)[0])
    assert borderline <= 0.01 * n * len(prim_idx) + 2
    assert (got["n_below"] > 0).sum() > len(prim_idx)  # the threshold actually screens something in


def test_proximity_screen_minutes_form_and_device_primaries():
    import torch
    batch, orc, jd, fr = _setup(4, 256, 150)
    mins = np.linspace(0.0, 4000.0, 150)
    pv, err = batch.propagate_from_epoch(mins)
    primaries = np.ascontiguousarray(pv[0:3, [3, 100], :].transpose(1, 0, 2))
    primaries[1, :, 10:20] = np.nan  # a primary without a state at some times is skipped there
    host = batch.propagate_proximity(mins, None, primaries, 2000.0)
    dev = batch.propagate_proximity(mins, None, torch.tensor(primaries, device="cuda:0"), 2000.0)
    assert np.array_equal(host.view(np.uint8), dev.view(np.uint8))
    pos = np.moveaxis(pv[0:3], 0, -1)
    d = np.linalg.norm(pos - np.moveaxis(primaries[1], 0, -1)[None], axis=-1)
    have = (err == 0) & np.isfinite(d)
    want = np.where(have, d, np.inf).min(1)
    assert np.allclose(host[:, 1]["d_min_km"], want, rtol=0, atol=1e-9)
    assert not np.isin(host[:, 1]["k_min"], np.arange(10, 20)).any()
