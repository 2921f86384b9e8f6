"""Multi-GPU host logic on CPU: satellite-range sharding, cost-balanced plans and the
world_size-2 gloo path (timing max-reduce, shard gather == unsharded result)."""
import os

import numpy as np
import pytest

import perturb_b200 as pb
from perturb_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 30000, 240001):
        for world in (1, 2, 3, 8):
            spans = [pb.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        pb.shard_range(10, 2, 2)


def test_shard_blocks_slices_the_catalogue_text():
    n = 101
    b1, b2 = pb.synth_catalog(3, n)
    parts = [pb.shard_blocks(b1, b2, n, r, 4) for r in range(4)]
    assert b"".join(p[0] for p in parts) == b1 and b"".join(p[1] for p in parts) == b2
    assert sum(p[2] for p in parts) == n


def test_estimate_class_matches_the_mix_of_each_config():
    tl, _ = pb.parse_tle_blocks(*pb.synth_catalog(4, 400), 400)
    cls = pb.estimate_class(tl)
    assert np.array_equal(cls[:200], np.full(200, 3)) and np.array_equal(cls[200:], np.full(200, 4))
    tl, _ = pb.parse_tle_blocks(*pb.synth_catalog(5, 300), 300)
    assert (pb.estimate_class(tl) == 0).all()
    tl, _ = pb.parse_tle_blocks(*pb.synth_catalog(3, 2000), 2000)
    counts = np.bincount(pb.estimate_class(tl), minlength=5)
    assert counts[2] == 300 and counts[3] == 200 and counts[4] == 100  # MEO+GTO, GEO, Molniya


def test_plan_ranges_with_device_shares():
    cost = np.ones(1000)
    assert sharding.plan_ranges(cost, 4, [1, 1, 1, 1]) == sharding.plan_ranges(cost, 4)
    r = sharding.plan_ranges(cost, 4, [20.0, 20.0, 20.0, 45.0])  # profiles/r2p_n4_e2e_nobind.json
    assert [hi - lo for lo, hi in r] == [190, 191, 190, 429] and r[-1][1] == 1000
    # shares weigh COST, not count: the deep-space half costs 1.8x per satellite
    cost = np.concatenate([np.full(500, 1170.0), np.full(500, 2097.0)])
    r = sharding.plan_ranges(cost, 2, [1.0, 3.0])
    got = [cost[lo:hi].sum() for lo, hi in r]
    assert abs(got[0] / sum(got) - 0.25) < 2e-3
    assert sharding.plan_ranges(np.ones(10), 3, [0.0, 1.0, 1.0])[0] == (0, 0)
    with pytest.raises(ValueError):
        sharding.plan_ranges(np.ones(10), 3, [1.0, 1.0])


def test_rate_ranges_are_proportional_contiguous_and_never_empty():
    from perturb_b200 import rate_ranges
    # the 8-GPU box of profiles/r2f_n8_e2e_nobind.json: four GPUs at 11.7 GB/s, four at 16.6
    r = rate_ranges(240000, [11.7] * 4 + [16.6] * 4)
    assert r[0][0] == 0 and r[-1][1] == 240000
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    counts = np.array([hi - lo for lo, hi in r])
    assert np.allclose(counts[:4], 240000 * 11.7 / (4 * 28.3), atol=1) and np.allclose(
        counts[4:], 240000 * 16.6 / (4 * 28.3), atol=1)
    assert rate_ranges(120000, [18.25] * 4) == [(i * 30000, (i + 1) * 30000) for i in range(4)]
    # a dead link keeps one tile; tiny catalogues are still partitioned exactly
    r = rate_ranges(1000, [0.0, 50.0, 50.0])
    assert r[0] == (0, 32) and r[-1][1] == 1000 and all(hi > lo for lo, hi in r)
    r = rate_ranges(5, [1.0, 1.0, 1.0])
    assert r[0][0] == 0 and r[-1][1] == 5 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    with pytest.raises(ValueError):
        rate_ranges(10, [])


def test_plan_ranges_balances_cost_not_count():
    rng = np.random.default_rng(0)
    # first half cheap, second half twice as expensive: the second shard must be shorter
    cost = np.concatenate([np.full(1000, 1170.0), np.full(1000, 2097.0)])
    ranges = pb.plan_ranges(cost, 2)
    assert ranges[0][0] == 0 and ranges[-1][1] == 2000 and ranges[0][1] == ranges[1][0]
    assert 1215 <= ranges[0][1] <= 1227  # 1000 cheap + 221 expensive = half the cost
    sums = [cost[l:h].sum() for l, h in ranges]
    assert abs(sums[0] - sums[1]) <= 2097.0
    for k in (1, 3, 8):
        cost = rng.choice(sharding.CLASS_COST, size=5000)
        ranges = pb.plan_ranges(cost, k)
        assert ranges[0][0] == 0 and ranges[-1][1] == 5000
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sums = np.array([cost[l:h].sum() for l, h in ranges])
        assert sums.max() - sums.min() <= 2 * 2097.0
    assert pb.plan_ranges(np.ones(2), 4)[-1][1] == 2


def _worker(rank, world, port, n, tmp):
    import torch.distributed as dist
    from oracle import port as oracle_port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b1, b2 = pb.synth_catalog(3, n)
        s1, s2, k = pb.shard_blocks(b1, b2, n, rank, world)
        L1, L2 = pb.block_to_lines(s1, k), pb.block_to_lines(s2, k)
        # the CPU oracle stands in for the device here: what is under test is the sharding,
        # the gather and the timing reduction, not the arithmetic
        orc = oracle_port.OracleSatellites.from_tle_lines(L1, L2, 1, "i", flavor=1)
        jd, fr = pb.synth_time_grid(16)
        err, pos, vel, _ = orc.propagate_grid(jd, fr, use_jd=True)
        lo, hi = pb.shard_range(n, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, err, pos))
        slowest = pb.max_over_ranks(10.0 + rank, dist)
        if rank == 0:
            np.savez(tmp, err=np.concatenate([g[2] for g in gathered]),
                     pos=np.concatenate([g[3] for g in gathered]),
                     spans=np.array([(g[0], g[1]) for g in gathered]), slowest=slowest)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_gather_to_the_unsharded_result(tmp_path):
    import torch.multiprocessing as mp
    from oracle import port as oracle_port
    n, world = 150, 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 1000, n, out), nprocs=world, join=True)
    got = np.load(out)
    assert got["spans"].tolist() == [[0, 75], [75, 150]]
    assert got["slowest"] == 11.0  # max over ranks, not rank 0's own 10.0
    b1, b2 = pb.synth_catalog(3, n)
    orc = oracle_port.OracleSatellites.from_tle_lines(
        pb.block_to_lines(b1, n), pb.block_to_lines(b2, n), 1, "i", flavor=1)
    jd, fr = pb.synth_time_grid(16)
    err, pos, _, _ = orc.propagate_grid(jd, fr, use_jd=True)
    assert np.array_equal(got["err"], err) and np.array_equal(got["pos"], pos, equal_nan=True)


@pytest.mark.gpu
def test_sharded_batch_over_available_devices():
    """In-process sharding (1..N visible devices): same bits as one unsharded batch."""
    import torch
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) * (2 if ndev == 1 else 1)  # two shards even on one GPU
    n = 900
    b1, b2 = pb.synth_catalog(3, n)
    L1, L2 = pb.block_to_lines(b1, n), pb.block_to_lines(b2, n)
    whole = pb.SatelliteBatch.from_tles(L1, L2)
    sharded = pb.ShardedSatelliteBatch.from_tles(L1, L2, devices)
    assert sum(h - l for l, h in sharded.ranges) == n
    jd, fr = pb.synth_time_grid(77)
    pa, ea = whole.propagate(jd, fr)
    ps, es = sharded.propagate(jd, fr)
    assert np.array_equal(ea, es) and np.array_equal(pa, ps, equal_nan=True)
    assert np.array_equal(whole.last_error(), sharded.last_error())
    assert np.array_equal(whole.orbit_class(), sharded.orbit_class())


@pytest.mark.gpu
def test_device_resident_gather_equals_the_host_gather():
    """Shards store their rows straight into ONE device's buffers (over NVLink peer access when
    they sit on other GPUs); the gathered planes equal the unsharded batch bit for bit, with a
    host or a foreign-device time grid, planes or StateVector records."""
    import torch
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) * (2 if ndev == 1 else 1)
    target = ndev - 1  # not the first shard's device when there are several
    n = 1100
    b1, b2 = pb.synth_catalog(3, n)
    L1, L2 = pb.block_to_lines(b1, n), pb.block_to_lines(b2, n)
    whole = pb.SatelliteBatch.from_tles(L1, L2, device=0)
    sharded = pb.ShardedSatelliteBatch.from_tles(L1, L2, devices)
    jd, fr = pb.synth_time_grid(130)
    pa, ea = whole.propagate(jd, fr)
    pd, ed = sharded.propagate_to_device(jd, fr, device=target)
    assert pd.device.index == target and ed.device.index == target
    assert np.array_equal(ea, ed.cpu().numpy())
    assert np.array_equal(pa, pd.cpu().numpy(), equal_nan=True)
    # the time grid itself living on the target device: copied to each shard's device
    d_jd = torch.tensor(jd, device="cuda:%d" % target)
    d_fr = torch.tensor(fr, device="cuda:%d" % target)
    pd2 = torch.full_like(pd, -1.0)
    ed2 = torch.full_like(ed, -1)
    torch.cuda.synchronize(target)
    sharded.propagate(d_jd, d_fr, pd2, ed2)
    assert torch.equal(ed, ed2) and np.array_equal(pd.cpu().numpy(), pd2.cpu().numpy(), equal_nan=True)
    # records mode through the same peer path: shard g of the last device writes into `target`
    last = sharded.batches[-1]
    lo, hi = sharded.ranges[-1]
    st = torch.empty((hi - lo, 130, 8), dtype=torch.float64, device="cuda:%d" % (0 if ndev > 1 else target))
    er = torch.empty((hi - lo, 130), dtype=torch.int32, device=st.device)
    torch.cuda.synchronize(st.device)
    last.propagate_states(jd, fr, st, er)
    assert np.array_equal(er.cpu().numpy(), ea[lo:hi])
    got = st.cpu().numpy()
    assert np.array_equal(np.moveaxis(got[:, :, 2:8], 2, 0), pa[:, lo:hi], equal_nan=True)
