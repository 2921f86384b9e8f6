"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/liboracle.so (sgp4_oracle.c), the
plain-C CPU restatement of the reference's SGP4 path. Importable by tests/,
__graft_entry__.smoke() and bench.py's CPU arms only -- never by perturb_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

_INT_FIELDS = "error isimp irez deep opsmode_a grav".split()
_DBL_FIELDS = (
    "aycof con41 cc1 cc4 cc5 d2 d3 d4 delmo eta argpdot omgcof sinmao t t2cof t3cof t4cof "
    "t5cof x1mth2 x7thm1 mdot nodedot xlcof xmcof nodecf "
    "d2201 d2211 d3210 d3222 d4410 d4422 d5220 d5232 d5421 d5433 dedt del1 del2 del3 didt "
    "dmdt dnodt domdt e3 ee2 peo pgho pho pinco plo se2 se3 sgh2 sgh3 sgh4 sh2 sh3 si2 si3 "
    "sl2 sl3 sl4 gsto xfact xgh2 xgh3 xgh4 xh2 xh3 xi2 xi3 xl2 xl3 xl4 xlamo zmol zmos "
    "atime xli xni "
    "a altp alta jdsatepoch jdsatepochF nddot ndot bstar inclo nodeo ecco argpo mo no_kozai "
    "no_unkozai am em im Om om mm nm tumin mus radiusearthkm xke j2 j3 j4 j3oj2 kepler_last cond"
).split()


class OrcSat(C.Structure):
    _fields_ = ([(n, C.c_int) for n in _INT_FIELDS] + [(n, C.c_double) for n in _DBL_FIELDS]
                + [("kepler_iters", C.c_int), ("pad_", C.c_int)])


class OrcTle(C.Structure):
    _fields_ = [
        (n, C.c_double)
        for n in (
            "epoch_year epoch_day_of_year n_dot n_ddot b_star inclination raan eccentricity "
            "arg_of_perigee mean_anomaly mean_motion"
        ).split()
    ]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or (
        os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "sgp4_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        assert L.orc_sizeof_sat() == C.sizeof(OrcSat), "orc_sat layout drift"
        L.orc_sat_from_tle.argtypes = [C.POINTER(OrcSat), C.POINTER(OrcTle), C.c_int, C.c_char]
        L.orc_sat_from_vallado_fields.argtypes = L.orc_sat_from_tle.argtypes
        L.orc_sgp4init.argtypes = [C.POINTER(OrcSat), C.c_int, C.c_char] + [C.c_double] * 10
        L.orc_parse_tle.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(OrcTle)]
        L.orc_propagate_mins.argtypes = [C.POINTER(OrcSat), C.c_double, dp, dp]
        L.orc_propagate_jd.argtypes = [C.POINTER(OrcSat), C.c_double, C.c_double, dp, dp]
        L.orc_propagate_grid.restype = C.c_double
        L.orc_propagate_grid.argtypes = [
            C.POINTER(OrcSat), C.c_long, dp, dp, C.c_long, C.c_int, C.c_int, dp, dp,
            C.POINTER(C.c_int)]
        L.orc_propagate_grid_diag.restype = C.c_double
        L.orc_propagate_grid_diag.argtypes = L.orc_propagate_grid.argtypes + [
            C.POINTER(C.c_ubyte), dp]
        L.orc_propagate_grid_elems.restype = C.c_double
        L.orc_propagate_grid_elems.argtypes = L.orc_propagate_grid.argtypes + [dp]
        L.orc_rv2coe.argtypes = [dp, dp, C.c_int, dp]
        L.orc_rv2coe.restype = None
        L.orc_gstime.restype = C.c_double
        L.orc_gstime.argtypes = [C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleSatellites:
    """An array of oracle satellite records (the restated `perturb::Satellite`)."""

    def __init__(self, n):
        self.n = int(n)
        self.sats = (OrcSat * self.n)()
        self.init_error = np.zeros(self.n, dtype=np.int32)
        self.parse_error = np.zeros(self.n, dtype=np.int32)

    @classmethod
    def from_tle_lines(cls, lines1, lines2, grav=1, opsmode="i", flavor=1):
        """flavor 1 = Satellite::from_tle route (Vallado's parser conventions);
        flavor 0 = TwoLineElement::parse + Satellite(TwoLineElement)."""
        self = cls(len(lines1))
        L = lib()
        self.tles = (OrcTle * self.n)()
        for i, (a, b) in enumerate(zip(lines1, lines2)):
            self.parse_error[i] = L.orc_parse_tle(
                a.encode(), b.encode(), flavor, C.byref(self.tles[i]))
        self._init_from_tles(grav, opsmode, flavor)
        return self

    def _init_from_tles(self, grav, opsmode, flavor):
        L = lib()
        self.grav, self.opsmode, self.flavor = grav, opsmode, flavor
        fn = L.orc_sat_from_vallado_fields if flavor == 1 else L.orc_sat_from_tle
        for i in range(self.n):
            fn(C.byref(self.sats[i]), C.byref(self.tles[i]), grav, opsmode.encode())
            self.init_error[i] = self.sats[i].error

    def perturbed(self, sign=+1):
        """A copy whose orbital elements are moved by one ulp each (alternating directions):
        the conditioning probe of tests/helpers.py."""
        other = OracleSatellites(self.n)
        other.tles = (OrcTle * self.n)()
        names = ("mean_motion eccentricity inclination raan arg_of_perigee mean_anomaly "
                 "b_star").split()
        for i in range(self.n):
            C.memmove(C.byref(other.tles[i]), C.byref(self.tles[i]), C.sizeof(OrcTle))
            for j, nm in enumerate(names):
                v = getattr(other.tles[i], nm)
                d = sign * (1.0 if j % 2 == 0 else -1.0)
                setattr(other.tles[i], nm, float(np.nextafter(v, v + d * (abs(v) + 1.0))))
        other._init_from_tles(self.grav, self.opsmode, self.flavor)
        return other

    @classmethod
    def from_raw(cls, el, grav=1, opsmode="i"):
        n = len(el["no_kozai"])
        self = cls(n)
        L = lib()
        for i in range(n):
            s = self.sats[i]
            s.jdsatepoch = float(el["jdsatepoch"][i])
            s.jdsatepochF = float(el["jdsatepochF"][i])
            epoch = (s.jdsatepoch + s.jdsatepochF) - 2433281.5
            L.orc_sgp4init(
                C.byref(s), grav, opsmode.encode(), epoch,
                *[float(el[k][i]) for k in
                  "bstar ndot nddot ecco argpo inclo mo no_kozai nodeo".split()])
            self.init_error[i] = s.error
        return self

    def field(self, name):
        return np.array([getattr(s, name) for s in self.sats])

    def epoch(self):
        return self.field("jdsatepoch"), self.field("jdsatepochF")

    def propagate_mins_one(self, i, mins):
        r = np.full(3, np.nan)
        v = np.full(3, np.nan)
        e = lib().orc_propagate_mins(C.byref(self.sats[i]), float(mins), _dp(r), _dp(v))
        return e, r, v

    def propagate_jd_one(self, i, jd, frac):
        r = np.full(3, np.nan)
        v = np.full(3, np.nan)
        e = lib().orc_propagate_jd(C.byref(self.sats[i]), float(jd), float(frac), _dp(r), _dp(v))
        return e, r, v

    def propagate_grid(self, ta, tb=None, use_jd=False, nthreads=1, want_state=True):
        ta = np.ascontiguousarray(ta, dtype=np.float64)
        tb = np.ascontiguousarray(tb if tb is not None else np.zeros_like(ta), dtype=np.float64)
        m = ta.size
        err = np.zeros((self.n, m), dtype=np.int32)
        if want_state:
            pos = np.zeros((self.n, m, 3))
            vel = np.zeros((self.n, m, 3))
            pp, vp = _dp(pos), _dp(vel)
        else:
            pos = vel = None
            pp = vp = None
        self.converged = np.ones((self.n, m), dtype=np.uint8)
        self.cond = np.zeros((self.n, m))
        secs = lib().orc_propagate_grid_diag(
            self.sats, self.n, _dp(ta), _dp(tb), m, int(bool(use_jd)), int(nthreads), pp, vp,
            err.ctypes.data_as(C.POINTER(C.c_int)),
            self.converged.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(self.cond))
        return err, pos, vel, secs

    def propagate_grid_elements(self, ta, tb=None, use_jd=False, nthreads=1):
        """(err [n,m], elems [n,m,7] = am, em, im, Om, om, mm, nm as sgp4() leaves them in the
        record; NaN where the call returned before storing them)."""
        ta = np.ascontiguousarray(ta, dtype=np.float64)
        tb = np.ascontiguousarray(tb if tb is not None else np.zeros_like(ta), dtype=np.float64)
        m = ta.size
        err = np.zeros((self.n, m), dtype=np.int32)
        elems = np.zeros((self.n, m, 7))
        lib().orc_propagate_grid_elems(
            self.sats, self.n, _dp(ta), _dp(tb), m, int(bool(use_jd)), int(nthreads), None, None,
            err.ctypes.data_as(C.POINTER(C.c_int)), _dp(elems))
        return err, elems


def rv2coe(r, v, grav=1):
    """orc_rv2coe over states r, v [..., 3] -> [..., 11] (ClassicalOrbitalElements order)."""
    r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((r.shape[0], 11))
    L = lib()
    for e in range(r.shape[0]):
        L.orc_rv2coe(_dp(r[e]), _dp(v[e]), grav, _dp(out[e]))
    return out
