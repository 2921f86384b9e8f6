"""perturb_b200 -- B200-native batched SGP4/SDP4 propagator (drop-in for the
`Satellite::propagate` path of gunvirranu/perturb). The product is the C-ABI shared library
`libperturb_gpu.so` (include/perturb_gpu.h); this package is its host-side Python mirror."""
from . import capi  # noqa: F401
from .batch import (  # noqa: F401
    COE_NAMES, SatelliteBatch, block_to_lines, classical_elements, parse_tle_blocks, parse_tles, synth_catalog,
    synth_time_grid)
from .sharding import (  # noqa: F401
    ShardedSatelliteBatch, estimate_class, max_over_ranks, plan_ranges, rate_ranges, shard_blocks,
    shard_range)
from .capi import WGS72, WGS72_OLD, WGS84, PerturbGpuError  # noqa: F401

SGP4_ERROR_NAMES = (
    "NONE", "MEAN_ELEMENTS", "MEAN_MOTION", "PERT_ELEMENTS", "SEMI_LATUS_RECTUM",
    "EPOCH_ELEMENTS_SUB_ORBITAL", "DECAYED", "INVALID_TLE", "UNKNOWN")
