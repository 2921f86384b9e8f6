"""Error bounds of the device FP64 fast paths (perturb_b200/csrc/sgp4_math.cuh) against
host values (numpy / glibc, < 1 ulp), over the argument ranges the SGP4 path produces."""
import ctypes as C

import numpy as np
import pytest

from perturb_b200 import capi

pytestmark = pytest.mark.gpu
N = 400000


def probe(fn, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.ones_like(x) if y is None else y, dtype=np.float64)
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    st = capi.tools().perturb_gpu_math_probe(
        fn, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), x.size)
    capi.check(st, "perturb_gpu_math_probe")
    return out


def ulps(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_sincos_reduced_angles():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-15.0, 15.0, N), rng.uniform(-1e-3, 1e-3, 1000),
                        np.linspace(-2 * np.pi, 2 * np.pi, 4001)])
    # absolute error relative to 1 (the quantity that propagates into positions)
    for fn, ref in ((0, np.sin), (1, np.cos), (2, np.sin), (3, np.cos)):
        err = np.abs(probe(fn, x) - ref(x))
        assert err.max() < 2.3e-16, (fn, err.max())
    small = rng.uniform(-0.7, 0.7, N)
    assert ulps(probe(0, small), np.sin(small)).max() <= 2.0
    assert ulps(probe(1, small), np.cos(small)).max() <= 2.0


def test_sincos_small_angle_path():
    rng = np.random.default_rng(2)
    d = np.concatenate([rng.uniform(-0.0156, 0.0156, N), rng.uniform(-3.0, 3.0, 5000), [0.0]])
    assert np.abs(probe(11, d) - np.sin(d)).max() < 2.3e-16
    assert np.abs(probe(12, d) - np.cos(d)).max() < 2.3e-16
    # the two-coefficient series (|d| < 2^-9: short-period corrections), same fall-back
    d9 = np.concatenate([rng.uniform(-0.00196, 0.00196, N), rng.uniform(-3.0, 3.0, 5000), [0.0]])
    assert np.abs(probe(15, d9) - np.sin(d9)).max() < 2.3e-16
    assert np.abs(probe(16, d9) - np.cos(d9)).max() < 2.3e-16
    tiny = rng.uniform(-0.0156, 0.0156, N)
    assert ulps(probe(11, tiny), np.sin(tiny)).max() <= 1.0


def test_division_reciprocal_sqrt():
    rng = np.random.default_rng(3)
    a = rng.uniform(-10.0, 10.0, N) * 10.0 ** rng.integers(-8, 8, N)
    b = rng.uniform(0.1, 10.0, N) * 10.0 ** rng.integers(-8, 8, N) * rng.choice([-1.0, 1.0], N)
    assert ulps(probe(4, a, b), a / b).max() <= 1.0
    assert ulps(probe(5, b), 1.0 / b).max() <= 1.5
    coarse = np.abs(probe(14, a, b) - a / b) / np.abs(a / b)
    assert coarse.max() < 1.0e-9  # seed + one Newton step; used for Newton corrections only
    p = np.abs(b)
    assert ulps(probe(6, p), np.sqrt(p)).max() <= 1.0
    assert ulps(probe(7, p), 1.0 / np.sqrt(p)).max() <= 2.0
    assert probe(6, np.zeros(4))[0] == 0.0


def test_angle_reductions():
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-5000.0, 5000.0, N), rng.uniform(-7.0, 7.0, 5000),
                        np.arange(-50, 51) * (2 * np.pi), [0.0, -0.0]])
    twopi = 2.0 * np.pi
    assert np.array_equal(probe(8, x), np.fmod(x, twopi))  # C fmod, bit for bit
    w = probe(9, x)
    assert (np.abs(w) <= np.pi * (1 + 1e-15)).all()
    k = np.rint((x - w) / twopi)
    assert np.abs(x - k * twopi - w).max() < 1e-12
    w2 = probe(10, x)
    # modulo the true 2 pi: sin/cos of the reduced angle match those of the original
    assert np.abs(np.sin(w2) - np.sin(x)).max() < 3e-16
    assert np.abs(np.cos(w2) - np.cos(x)).max() < 3e-16


def test_pow_two_thirds():
    rng = np.random.default_rng(5)
    x = rng.uniform(0.5, 40.0, N)
    assert ulps(probe(13, x), np.power(x, 2.0 / 3.0)).max() <= 3.0
