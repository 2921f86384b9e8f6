"""GPU tests of the drop-in boundary added in round 2: host outputs in ordinary (pageable)
memory through the bounce pipeline, `perturb_gpu_propagate_into` with the reference's
untouched-on-error semantics, per-handle ordering of calls on different streams, the write-back
of `sat_rec` (N2), argument validation of the Python mirror."""
import json
import os
import subprocess

import numpy as np
import pytest

import perturb_b200 as pb
from helpers import load_ver_golden, ver_lines
from perturb_b200 import capi

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = load_ver_golden()


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


def _mixed_batch(n):
    b1, b2 = pb.synth_catalog(3, n)
    return pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", n=n)


def test_pageable_pinned_and_device_outputs_carry_the_same_bits():
    import torch
    n, m = 300, 200
    batch = _mixed_batch(n)
    jd, fr = pb.synth_time_grid(m)
    dev = torch.device("cuda", 0)
    d_out = torch.empty((6, n, m), dtype=torch.float64, device=dev)
    d_err = torch.empty((n, m), dtype=torch.int32, device=dev)
    batch.propagate(torch.tensor(jd, device=dev), torch.tensor(fr, device=dev), d_out, d_err)
    want, want_err = d_out.cpu().numpy(), d_err.cpu().numpy()
    # ordinary numpy memory: pinned bounce buffers + host threads
    got, got_err = batch.propagate(jd, fr)
    assert np.array_equal(_bits(got), _bits(want)) and np.array_equal(got_err, want_err)
    # a wider caller matrix filled one time-chunk at a time (row_stride > m), pageable
    wide = np.full((6, n, 3 * m), -7.0)
    wide_err = np.full((n, 3 * m), -7, dtype=np.int32)
    batch.propagate(jd, fr, wide[:, :, m:2 * m], wide_err[:, m:2 * m])
    assert np.array_equal(_bits(wide[:, :, m:2 * m]), _bits(want))
    assert np.array_equal(wide_err[:, m:2 * m], want_err)
    assert (wide[:, :, :m] == -7.0).all() and (wide[:, :, 2 * m:] == -7.0).all()
    assert (wide_err[:, :m] == -7).all() and (wide_err[:, 2 * m:] == -7).all()
    # pinned memory: written by the copy engine directly
    p_out = torch.empty((6, n, m), dtype=torch.float64).pin_memory()
    p_err = torch.empty((n, m), dtype=torch.int32).pin_memory()
    batch.propagate(jd, fr, p_out, p_err)
    assert np.array_equal(_bits(p_out.numpy()), _bits(want)) and np.array_equal(p_err.numpy(), want_err)
    # StateVector records, pageable and pinned
    rec, rec_err = batch.propagate_states(jd, fr)
    assert np.array_equal(rec_err, want_err)
    assert np.array_equal(_bits(rec["position"][..., 0]), _bits(want[0]))
    assert np.array_equal(_bits(rec["velocity"][..., 2]), _bits(want[5]))
    p_rec = torch.empty((n, m, 8), dtype=torch.float64).pin_memory()
    batch.propagate_states(jd, fr, p_rec, p_err)
    assert np.array_equal(_bits(p_rec.numpy().reshape(n, m, 8)),
                          _bits(rec.view(np.float64).reshape(n, m, 8)))


def test_pageable_output_larger_than_one_bounce_chunk():
    """3000 x 1100 records = 224 MB: two chunks of the 128 MiB bounce buffers, odd chunk tails."""
    import torch
    n, m = 3000, 1101
    batch = _mixed_batch(n)
    jd, fr = pb.synth_time_grid(m)
    dev = torch.device("cuda", 0)
    d_rec = torch.empty((n, m, 8), dtype=torch.float64, device=dev)
    d_err = torch.empty((n, m), dtype=torch.int32, device=dev)
    batch.propagate_states(torch.tensor(jd, device=dev), torch.tensor(fr, device=dev), d_rec, d_err)
    rec, err = batch.propagate_states(jd, fr)
    assert np.array_equal(err, d_err.cpu().numpy())
    assert np.array_equal(_bits(rec.view(np.float64).reshape(n, m, 8)), _bits(d_rec.cpu().numpy()))
    out, err2 = batch.propagate(jd, fr)
    assert np.array_equal(err2, err)
    assert np.array_equal(_bits(out[1]), _bits(rec["position"][..., 1]))
    assert np.array_equal(_bits(out[4]), _bits(rec["velocity"][..., 1]))


@pytest.mark.parametrize("ops", ["a", "i"])
def test_propagate_into_leaves_failed_states_untouched(ops):
    """The reference returns before storing r, v for codes 1-4 (sgp4.cpp:1865,1878,1943,1995):
    the caller's StateVector keeps what it held; sv.epoch is always set (perturb.cpp:229,240)."""
    l1, l2 = ver_lines(G)
    batch = pb.SatelliteBatch.from_tles(l1, l2, pb.WGS72, ops, flavor=1)
    mins = np.concatenate([np.arange(0.0, 4320.0, 120.0), [494.2, 1560.0, 25.0, 55.0, 440.0]])
    nan_rec, nan_err = batch.propagate_states(mins)  # NaN convention of the C ABI
    states = np.zeros((batch.n, mins.size), dtype=batch.STATE_DTYPE)
    states["position"] = -1.0
    states["velocity"] = -2.0
    states["epoch_jd"] = -3.0
    err = np.full((batch.n, mins.size), -9, dtype=np.int32)
    batch.propagate_into(mins, None, states, err)
    assert np.array_equal(err, nan_err)
    failed = (err >= 1) & (err <= 4)
    assert failed.any() and (err == 6).any()  # the catalogue has codes 1, (2,) 4 and 6
    assert (states["position"][failed] == -1.0).all() and (states["velocity"][failed] == -2.0).all()
    ok = ~failed
    assert np.array_equal(_bits(states["position"][ok]), _bits(nan_rec["position"][ok]))
    assert np.array_equal(_bits(states["velocity"][ok]), _bits(nan_rec["velocity"][ok]))
    assert np.array_equal(_bits(states["epoch_jd"]), _bits(nan_rec["epoch_jd"]))
    assert np.array_equal(_bits(states["epoch_jd_frac"]), _bits(nan_rec["epoch_jd_frac"]))
    # the same into memory the caller page-locked in place: direct DMA + save / restore
    pinned = np.zeros_like(states)
    pinned["position"] = -1.0
    pinned["velocity"] = -2.0
    perr = np.full_like(err, -9)
    capi.check(capi.lib().perturb_gpu_host_register(pinned.ctypes.data, pinned.nbytes), "host_register")
    try:
        batch.propagate_into(mins, None, pinned, perr)
    finally:
        capi.lib().perturb_gpu_host_unregister(pinned.ctypes.data)
    assert np.array_equal(perr, err)
    assert np.array_equal(pinned.view(np.int64), states.view(np.int64))
    # err may be omitted, like a caller who ignores the return value
    again = np.zeros_like(states)
    again["position"] = -1.0
    again["velocity"] = -2.0
    batch.propagate_into(mins, None, again, None)
    assert np.array_equal(again.view(np.int64), states.view(np.int64))


def test_calls_on_different_streams_are_ordered_per_handle():
    """Scratch (time grid, resonance node table) belongs to the handle: a second call on another
    stream is ordered after the first instead of overwriting what its kernels still read."""
    import torch
    n, m = 4000, 640
    b1, b2 = pb.synth_catalog(4, n)  # resonant satellites: node table in use
    batch = pb.SatelliteBatch.from_text(b1, b2, pb.WGS72, "i", n=n)
    jd, fr = pb.synth_time_grid(2 * m)
    dev = torch.device("cuda", 0)
    grids = [(jd[:m], fr[:m]), (jd[m:] + 11.0, fr[m:])]  # host grids: both go through d_times
    want = [batch.propagate(a, b) for a, b in grids]
    outs = [(torch.empty((6, n, m), dtype=torch.float64, device=dev),
             torch.empty((n, m), dtype=torch.int32, device=dev)) for _ in grids]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    torch.cuda.synchronize(dev)
    for rep in range(3):
        for (a, b), (o, e), s in zip(grids, outs, streams):
            batch.propagate(a, b, o, e, stream=s.cuda_stream, sync=False)
    torch.cuda.synchronize(dev)
    for (o, e), (wo, we) in zip(outs, want):
        assert np.array_equal(e.cpu().numpy(), we)
        assert np.array_equal(_bits(o.cpu().numpy()), _bits(wo))


def test_mean_elements_into_host_buffers_equal_the_device_ones():
    n, m = 700, 333
    batch = _mixed_batch(n)
    mins = np.linspace(-300.0, 9000.0, m)
    pv, err, el = batch.propagate_elements(mins)
    hpv, herr, hel = batch.propagate_elements(mins, host=True)
    assert np.array_equal(herr, err.cpu().numpy())
    assert np.array_equal(_bits(hpv), _bits(pv.cpu().numpy()))
    assert np.array_equal(_bits(hel), _bits(el.cpu().numpy()))


def test_exported_records_agree_with_the_field_queries():
    n = 500
    batch = _mixed_batch(n)
    rec = batch.export_records()
    assert rec.shape == (n, 1000)
    layout = {}
    for ln in open(os.path.join(ROOT, "include", "perturb_elsetrec_layout.h")):
        ln = ln.strip()
        if ln.startswith("X("):
            name, off, kind = [x.strip() for x in ln[2:ln.index(")")].split(",")]
            layout[name] = (int(off), kind)
    for name in ("no_unkozai", "cc1", "xlamo", "d5433", "jdsatepochF", "am", "gsto", "j3oj2"):
        off = layout[name][0]
        got = rec[:, off:off + 8].copy().view(np.float64)[:, 0]
        assert np.array_equal(_bits(got), _bits(batch.record_field(name))), name
    irez = rec[:, layout["irez"][0]:layout["irez"][0] + 4].copy().view(np.int32)[:, 0]
    assert np.array_equal(irez, batch.record_field("irez").astype(np.int32))
    method = rec[:, layout["method"][0]]
    assert np.array_equal(method == ord("d"), batch.orbit_class() >= 2)
    assert (rec[:, layout["init"][0]] == ord("n")).all()
    assert (rec[:, layout["operationmode"][0]] == ord("i")).all()
    # write_back of minute 0 reproduces what the closing sgp4(t = 0) of sgp4init stored
    before = rec.copy()
    batch.write_back(rec, 0.0)
    for name in ("am", "em", "im", "nm"):
        off = layout[name][0]
        a = rec[:, off:off + 8].copy().view(np.float64)[:, 0]
        b = before[:, off:off + 8].copy().view(np.float64)[:, 0]
        assert np.allclose(a, b, rtol=1e-13, atol=0.0), name
    off = layout["t"][0]
    assert (rec[:, off:off + 8].copy().view(np.float64) == 0.0).all()


def test_write_back_matches_the_reference_records():
    """tests/cpp/writeback_check.cpp (built by __graft_entry__.build() against the reference's
    headers, linked with oracle/_ref): sat_rec after export_records / write_back against the
    reference's own sat_rec after the same call sequence."""
    exe = os.path.join(ROOT, "perturb_b200", "bin", "writeback_check")
    if not os.path.exists(exe) or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libperturb_ref.so")):
        pytest.skip("writeback_check was not built (needs the reference tree at build time)")
    for cfg, n in ((3, 3000), (4, 1000)):
        out = subprocess.run([exe, str(cfg), str(n)], capture_output=True, text=True, timeout=600)
        print(out.stdout, out.stderr[-3000:])
        assert out.returncode == 0, out.stdout + out.stderr[-2000:]
        r = json.loads(out.stdout.strip().splitlines()[-1])
        print(json.dumps(r))
        assert r["init_error_mismatch"] == 0
        assert r["export"]["int_char_mismatch"] == 0 and r["write_back"]["int_char_mismatch"] == 0
        assert r["write_back"]["code_mismatch"] == 0
        # worst difference as a fraction of the allowance (1e-11 |x| + 1e-14 for the initialised
        # record, 1e-9 |x| + 1e-10 for the per-call members; t and atime bit for bit)
        assert r["export"]["worst_ratio"] < 1.0, r["export"]
        assert r["write_back"]["worst_ratio"] < 1.0, r["write_back"]


def test_argument_validation_of_the_python_mirror():
    import torch
    batch = _mixed_batch(64)
    jd, fr = pb.synth_time_grid(10)
    dev = torch.device("cuda", 0)
    with pytest.raises(TypeError):
        batch.propagate(torch.tensor(jd, device=dev).float(), torch.tensor(fr, device=dev))
    with pytest.raises(ValueError):
        batch.propagate(jd, fr[:5])
    with pytest.raises(TypeError):
        batch.propagate(jd, fr, np.zeros((6, 64, 10), dtype=np.float32))
    with pytest.raises(ValueError):
        batch.propagate(jd, fr, np.zeros((6, 64, 9)))
    with pytest.raises(ValueError):
        batch.propagate(jd, fr, np.zeros((6, 64, 12))[:, :, :10], np.zeros((64, 10), dtype=np.int32))
    with pytest.raises(ValueError):
        batch.propagate(jd, fr, np.zeros((6, 64, 20))[:, :, ::2])
    with pytest.raises(ValueError):
        pb.SatelliteBatch.from_text(b"x" * 100, b"y" * 100, n=5)
    # integer time arrays are converted, not reinterpreted
    out_a, err_a = batch.propagate_from_epoch(np.arange(10))
    out_b, err_b = batch.propagate_from_epoch(np.arange(10, dtype=np.float64))
    assert np.array_equal(_bits(out_a), _bits(out_b)) and np.array_equal(err_a, err_b)
    assert capi.lib().perturb_gpu_elsetrec_size() == 1000
