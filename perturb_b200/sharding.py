"""Host-side sharding of a catalogue over GPUs.

The path shards with no exchange step: every (satellite, time) evaluation is independent and
the time grid is replicated, so a shard is a contiguous SATELLITE RANGE and there is no
data-path collective. Two users:

* `bench.py` under torchrun (one process per GPU): `shard_range` / `shard_blocks` pick this
  rank's range of the global catalogue, `max_over_ranks` is the only collective (timings);
* `ShardedSatelliteBatch` (one process, several devices): ranges balanced by estimated cost,
  every device copying its rows straight into its slice of one host buffer.
"""
import math

import numpy as np

from . import capi
from .batch import SatelliteBatch, parse_tles

# flops-v0 weights per orbit class: near full / near simplified /
# deep non-resonant / 24 h resonant / 12 h resonant
CLASS_COST = (1170.0, 1090.0, 1502.0, 1586.0, 2097.0)


def shard_range(n, rank, world):
    """Equal-count contiguous split: [lo, hi) of rank `rank` out of `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_blocks(block1, block2, n_total, rank, world, stride=capi.SYNTH_LINE_STRIDE):
    """This rank's slice of fixed-stride TLE line blocks. Returns (b1, b2, n_local)."""
    lo, hi = shard_range(n_total, rank, world)
    return block1[lo * stride: hi * stride], block2[lo * stride: hi * stride], hi - lo


def estimate_class(tles):
    """Orbit class per TLE from its numbers alone (host, no GPU): the same thresholds as
    sgp4init (src/sgp4.cpp:1590, 773-776, 1516) applied to the Kozai mean motion, which is
    within ~0.1 % of the un-Kozai'd one -- good enough to balance work, never used for results."""
    n = len(tles)
    out = np.zeros(n, dtype=np.uint8)
    for i, t in enumerate(tles):
        no = t.mean_motion * 2.0 * math.pi / 1440.0  # rad/min
        if not (no > 0.0):
            continue
        if 2.0 * math.pi / no >= 225.0:
            if 0.0034906585 < no < 0.0052359877:
                out[i] = 3
            elif 8.26e-3 <= no <= 9.24e-3 and t.eccentricity >= 0.5:
                out[i] = 4
            else:
                out[i] = 2
        else:
            a = (0.0743669161 / no) ** (2.0 / 3.0)
            out[i] = 1 if a * (1.0 - t.eccentricity) < 220.0 / 6378.135 + 1.0 else 0
    return out


def plan_ranges(costs, nshards, shares=None):
    """Contiguous ranges [(lo, hi)] * nshards whose summed cost is as equal -- or, with `shares`
    (one non-negative number per shard), as proportional to them -- as a prefix split allows
    (every satellite lands in exactly one range; empty ranges only if n < nshards or a share
    is zero)."""
    costs = np.asarray(costs, dtype=np.float64)
    n = costs.size
    if nshards < 1:
        raise ValueError("nshards must be >= 1")
    frac = np.arange(nshards + 1, dtype=np.float64) / nshards
    if shares is not None:
        sh = np.maximum(np.asarray(shares, dtype=np.float64), 0.0)
        if sh.size != nshards:
            raise ValueError("one share per shard")
        if sh.sum() > 0.0:
            frac = np.concatenate([[0.0], np.cumsum(sh) / sh.sum()])
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    total = cum[-1]
    cuts = [0]
    for s in range(1, nshards):
        target = total * frac[s]
        k = int(np.searchsorted(cum, target, side="left"))
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nshards)]


def rate_ranges(n, rates, min_count=32):
    """Contiguous ranges [(lo, hi)] whose satellite COUNTS are proportional to `rates` (e.g. the
    D2H rate each GPU of a box reaches while all of them copy: host-gathered output is bound by
    the links, not by the kernels, and the GPUs of a box do not always share them evenly). Every
    range keeps at least `min_count` satellites (one tile) while n allows it."""
    rates = np.maximum(np.asarray(rates, dtype=np.float64), 1e-9)
    k = rates.size
    if k < 1:
        raise ValueError("rates must name at least one shard")
    min_count = min(int(min_count), n // k)
    cuts = np.round(np.concatenate([[0.0], np.cumsum(rates / rates.sum())]) * n).astype(np.int64)
    cuts[0], cuts[-1] = 0, n
    for r in range(1, k):
        cuts[r] = min(max(cuts[r], cuts[r - 1] + min_count), n - min_count * (k - r))
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(k)]


def max_over_ranks(value, dist=None, device=None):
    """MAX-reduce a per-rank timing; identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class ShardedSatelliteBatch:
    """One process, several GPUs: satellite ranges, cost balanced, gathered to host memory."""

    def __init__(self, batches, ranges, n):
        self.batches, self.ranges, self.n = batches, ranges, n

    @classmethod
    def from_parsed(cls, tles, devices, grav=capi.WGS72, opsmode="i", shares=None):
        """`shares`: relative share of the cost per device instead of equal ones, e.g. each GPU's
        device -> host copy rate when the results are gathered into host memory."""
        n = len(tles)
        cost = np.array(CLASS_COST)[estimate_class(tles)]
        ranges = plan_ranges(cost, len(devices), shares)
        batches = []
        for dev, (lo, hi) in zip(devices, ranges):
            sub = (capi.Tle * (hi - lo))(*tles[lo:hi])
            batches.append(SatelliteBatch.from_parsed(sub, grav, opsmode, device=dev))
        return cls(batches, ranges, n)

    @classmethod
    def from_tles(cls, lines1, lines2, devices, grav=capi.WGS72, opsmode="i", flavor=1, shares=None):
        tles, _ = parse_tles(lines1, lines2, flavor)
        return cls.from_parsed(tles, devices, grav, opsmode, shares)

    def last_error(self):
        return np.concatenate([b.last_error() for b in self.batches]) if self.n else np.zeros(0, np.int32)

    def orbit_class(self):
        return np.concatenate([b.orbit_class() for b in self.batches]) if self.n else np.zeros(0, np.uint8)

    def propagate(self, jd, jd_frac, posvel=None, err=None):
        """All shards enqueue first (each into its row range of the shared buffers), then all
        are synchronised: devices run and copy concurrently. `posvel` / `err` may be host arrays
        (numpy / pinned torch: every device copies its rows in) or CUDA torch tensors on ONE
        device (device-resident gather: every shard's kernel stores its rows straight into
        that device's memory over NVLink peer access, no staging and no collective)."""
        m = int(jd.shape[0])
        if posvel is None:
            posvel = np.empty((6, self.n, m), dtype=np.float64)
        if err is None:
            err = np.empty((self.n, m), dtype=np.int32)
        for b, (lo, hi) in zip(self.batches, self.ranges):
            if hi > lo:
                b.propagate(jd, jd_frac, posvel[:, lo:hi, :], err[lo:hi, :], sync=False)
        for b in self.batches:
            b.synchronize()
        return posvel, err

    def propagate_to_device(self, jd, jd_frac, device=0):
        """Device-resident gather: returns (posvel [6, n, m], err [n, m]) as CUDA torch tensors
        on `device`, every shard having written its own rows there while it computed."""
        import torch
        dev = torch.device("cuda", device)
        m = int(jd.shape[0])
        posvel = torch.empty((6, self.n, m), dtype=torch.float64, device=dev)
        err = torch.empty((self.n, m), dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)  # the allocator's stream has nothing to do with the shards'
        return self.propagate(jd, jd_frac, posvel, err)
