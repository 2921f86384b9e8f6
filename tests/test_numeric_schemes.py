"""CPU checks of the numerical schemes kernel 2 uses in place of the reference's iterations
and library calls (perturb_b200/csrc/sgp4_eval.cuh). Each scheme is restated in numpy, operation
for operation, and held against the quantity the reference computes: Kepler's loop
(src/sgp4.cpp:1965-1983), sqrt(1 - el2) (:1998), the solar / lunar mean anomalies of dpper
(:262, :286). The CUDA code itself is tested on the GPU (tests/test_gpu_parity.py); these tests
pin the error bounds DESIGN.md quotes for the formulas."""
import numpy as np


def _reference_kepler(u, axnl, aynl):
    """The reference's loop: Newton from E = u, corrections clamped to +-0.95, at most 10, stop
    after one below 1e-12 (not applied). Returns sin E, cos E, passes, converged."""
    eo1 = u.copy()
    s, c = np.sin(eo1), np.cos(eo1)
    active = np.ones(u.shape, dtype=bool)
    passes = np.zeros(u.shape, dtype=int)
    for _ in range(10):
        sn, cn = np.sin(eo1), np.cos(eo1)
        s, c = np.where(active, sn, s), np.where(active, cn, c)
        tem5 = (u - aynl * cn + axnl * sn - eo1) / (1.0 - cn * axnl - sn * aynl)
        tem5 = np.clip(tem5, -0.95, 0.95)
        passes += active
        eo1 = np.where(active, eo1 + tem5, eo1)
        active = active & (np.abs(tem5) >= 1e-12)
    return s, c, passes, ~active


def _root(u, axnl, aynl):
    e = u.astype(np.longdouble)
    ax, ay, uu = axnl.astype(np.longdouble), aynl.astype(np.longdouble), u.astype(np.longdouble)
    e = e + (ax * np.sin(e) - ay * np.cos(e))  # one fixed-point step as a start
    for _ in range(12):
        f = e - uu - (ax * np.sin(e) - ay * np.cos(e))
        e = e - f / (1.0 - (ax * np.cos(e) + ay * np.sin(e)))
    return e


def _samples(e, n, rng):
    u = rng.uniform(-np.pi, np.pi, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    # half of the samples around the perigee (E - phi small), where the derivative is 1 - e
    k = n // 2
    u[:k] = phi[:k] + rng.normal(0.0, 0.05, k) * rng.choice([1.0, 0.1, 0.01], k)
    return u, e * np.cos(phi), e * np.sin(phi)


def kepler_small(u, axnl, aynl):
    """sgp4_eval.cuh kepler_small_sc: fourth-order solution for d = E - u, one Newton step with
    Taylor sums for the small d, (sin E, cos E) = (sin u, cos u) rotated by d."""
    su, cu = np.sin(u), np.cos(u)
    g0 = axnl * su - aynl * cu
    g1 = axnl * cu + aynl * su
    a = ((g1 * g1 + g1) * g1 + g1) + 1.0
    b = g1 * (5.0 / 3.0) + 0.5
    d = g0 * (a - g0 * g0 * b)
    z = d * d
    sd = d + d * z * (-1.0 / 6.0 + z / 120.0)
    cd = 1.0 + z * (-0.5 + z * (1.0 / 24.0 - z / 720.0))
    f = d - g0 * cd - g1 * sd
    w = g1 * cd - g0 * sd
    step = -f * ((w * w + w) + 1.0)
    s2 = sd + cd * step
    c2 = cd - sd * step
    return su * c2 + cu * s2, cu * c2 - su * s2


def kepler_halley(u, axnl, aynl):
    """sgp4_eval.cuh kepler_halley: Halley's iteration from E = u (Newton where the Halley
    factor is below 1/2), stop once a step is below 2^-10, rotate by it, one Newton correction
    from the rotated values. Returns sin E, cos E, passes with a general sin/cos, found."""
    eo1 = u.copy()
    done = np.zeros(u.shape, dtype=bool)
    passes = np.zeros(u.shape, dtype=int)
    se, ce, d = np.sin(eo1), np.cos(eo1), np.zeros(u.shape)
    for _ in range(8):
        se, ce = np.sin(eo1), np.cos(eo1)
        g = axnl * se - aynl * ce
        fp = 1.0 - (axnl * ce + aynl * se)
        f = (eo1 - u) - g
        dn = -f / fp
        q = 1.0 + 0.5 * dn * g / fp
        d = np.where(q > 0.5, dn / q, dn)
        d = np.clip(d, -0.95, 0.95)
        passes += ~done
        done = done | (np.abs(d) < 0.0009765625)
        eo1 = np.where(done, eo1, eo1 + d)
        if done.all():
            break
    z = d * d
    sd = d + d * z * (-1.0 / 6.0 + z / 120.0)
    cd = 1.0 + z * (-0.5 + z * (1.0 / 24.0 - z / 720.0))
    s1 = ce * sd + se * cd
    c1 = ce * cd - se * sd
    e1 = eo1 + d
    d1 = -((e1 - u) - (axnl * s1 - aynl * c1)) / (1.0 - (axnl * c1 + aynl * s1))
    return s1 + c1 * d1, c1 - s1 * d1, passes, done


def test_closed_form_kepler_for_nearly_circular_orbits():
    rng = np.random.default_rng(5)
    for e in (1e-6, 1e-3, 5e-3, 0.0199):  # the device takes this path for e^2 < 4e-4
        u, ax, ay = _samples(e, 60000, rng)
        s, c = kepler_small(u, ax, ay)
        root = _root(u, ax, ay)
        err = np.maximum(np.abs(s - np.sin(root)), np.abs(c - np.cos(root))).astype(np.float64)
        assert err.max() < 6e-15, (e, err.max())
        # the reference's own answer is within 1e-12 of the same root
        sr, cr, _, conv = _reference_kepler(u, ax, ay)
        assert conv.all()
        assert np.abs(s - sr).max() < 2e-12 and np.abs(c - cr).max() < 2e-12


def test_third_order_kepler_iteration_for_eccentric_orbits():
    rng = np.random.default_rng(6)
    for e in (0.02, 0.1, 0.3, 0.5, 0.7, 0.74, 0.7999):  # the device takes it for e^2 < 0.64
        u, ax, ay = _samples(e, 60000, rng)
        s, c, passes, found = kepler_halley(u, ax, ay)
        assert found.all() and passes.max() <= 3, (e, passes.max())
        root = _root(u, ax, ay)
        err = np.maximum(np.abs(s - np.sin(root)), np.abs(c - np.cos(root))).astype(np.float64)
        assert err.max() < 1e-14, (e, err.max())
        sr, cr, ref_passes, conv = _reference_kepler(u, ax, ay)
        assert conv.all()  # below e = 0.8 the reference's loop always converges within 10 passes
        assert np.abs(s - sr).max() < 6e-12 and np.abs(c - cr).max() < 6e-12
        if e >= 0.5:
            assert passes.mean() + 1.5 < ref_passes.mean(), (e, passes.mean(), ref_passes.mean())


def test_series_for_sqrt_one_minus_small():
    x = np.concatenate([np.linspace(0.0, 4e-4, 100001), [1e-300, 1e-12]])
    p = -0.0390625 * x - 0.0625
    p = p * x - 0.125
    p = p * x - 0.5
    got = p * x + 1.0
    want = np.sqrt(1.0 - x.astype(np.longdouble)).astype(np.float64)
    assert np.abs(got - want).max() < 2.3e-16


def test_solar_lunar_phase_split_reproduces_the_angles():
    """ThirdBodyPhase: zm = zmos + zns t with t = (jd - jd0) * 1440 split into a per-satellite part
    (zmos - zns * minutes(jd0)) and a per-time part (zns * minutes(jd)), minutes counted from one
    fixed reference date: the recombined sin / cos equal those of the reference's angle."""
    rng = np.random.default_rng(7)
    ref_jd = 2451545.0
    for rate in (1.19459e-5, 1.5835218e-4):  # zns, znl [rad/min]
        jd0 = 2451545.0 + rng.uniform(-8000.0, 11000.0, 20000).round(0) + 0.5
        fr0 = rng.uniform(0.0, 1.0, 20000)
        zmo = rng.uniform(0.0, 2.0 * np.pi, 20000)
        jd = jd0 + rng.integers(-30, 400, 20000)
        fr = rng.uniform(0.0, 1.0, 20000)
        t = ((jd - jd0) + (fr - fr0)) * 1440.0  # perturb.cpp:80-83, 237-238
        want = (zmo.astype(np.longdouble) + np.longdouble(rate) * t.astype(np.longdouble))
        sat = zmo - rate * (((jd0 - ref_jd) + fr0) * 1440.0)
        tim = rate * (((jd - ref_jd) + fr) * 1440.0)
        s = np.sin(sat) * np.cos(tim) + np.cos(sat) * np.sin(tim)
        c = np.cos(sat) * np.cos(tim) - np.sin(sat) * np.sin(tim)
        err = np.maximum(np.abs(s - np.sin(want)), np.abs(c - np.cos(want))).astype(np.float64)
        assert err.max() < 2e-12, (rate, err.max())  # a few ulp of angles of ~2e3 rad
