"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_ref/libperturb_ref.so.

That library is the UNMODIFIED reference (gunvirranu/perturb) compiled from
/root/reference by oracle/Makefile; this module just drives it. It may be imported by
tests/, __graft_entry__.smoke() and bench.py's CPU arms only -- never by perturb_b200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libperturb_ref.so")

# Order of ref_dump_fields() in oracle/ref_driver.cpp.
FIELD_NAMES = (
    "error isimp irez method_d "
    "aycof con41 cc1 cc4 cc5 d2 d3 d4 delmo eta argpdot omgcof sinmao t2cof t3cof t4cof t5cof "
    "x1mth2 x7thm1 mdot nodedot xlcof xmcof nodecf "
    "d2201 d2211 d3210 d3222 d4410 d4422 d5220 d5232 d5421 d5433 dedt del1 del2 del3 didt "
    "dmdt dnodt domdt e3 ee2 peo pgho pho pinco plo se2 se3 sgh2 sgh3 sgh4 sh2 sh3 si2 si3 "
    "sl2 sl3 sl4 gsto xfact xgh2 xgh3 xgh4 xh2 xh3 xi2 xi3 xl2 xl3 xl4 xlamo zmol zmos "
    "atime xli xni a altp alta jdsatepoch jdsatepochF nddot ndot bstar inclo nodeo ecco "
    "argpo mo no_kozai no_unkozai am em im Om om mm nm "
    "tumin mus radiusearthkm xke j2 j3 j4 j3oj2"
).split()


class RefTle(C.Structure):
    _fields_ = [("catalog_number", C.c_char * 8)] + [
        (n, C.c_double)
        for n in (
            "epoch_year epoch_day_of_year n_dot n_ddot b_star inclination raan eccentricity "
            "arg_of_perigee mean_anomaly mean_motion revolution_number element_set_number "
            "ephemeris_type classification line_1_checksum line_2_checksum"
        ).split()
    ]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.ref_alloc.restype = C.c_void_p
        L.ref_alloc.argtypes = [C.c_long]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_init_from_tle.argtypes = [C.c_void_p, C.c_long, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_parse_tle.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(RefTle)]
        L.ref_init_from_parsed.argtypes = [C.c_void_p, C.c_long, C.POINTER(RefTle), C.c_int]
        L.ref_init_verif.argtypes = [
            C.c_void_p, C.c_long, C.c_char_p, C.c_char_p, C.c_char, C.c_int, dp, dp, dp]
        L.ref_init_raw.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_char] + [C.c_double] * 11
        L.ref_last_error.argtypes = [C.c_void_p, C.c_long]
        L.ref_epoch.argtypes = [C.c_void_p, C.c_long, dp, dp]
        L.ref_propagate_mins.argtypes = [C.c_void_p, C.c_long, C.c_double, dp, dp]
        L.ref_propagate_jd.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_double, dp, dp]
        L.ref_dump_fields.argtypes = [C.c_void_p, C.c_long, dp]
        L.ref_propagate_grid.restype = C.c_double
        L.ref_propagate_grid.argtypes = [
            C.c_void_p, C.c_long, dp, dp, C.c_long, C.c_int, C.c_int, dp, dp,
            C.POINTER(C.c_int)]
        L.ref_rv2coe.argtypes = [dp, dp, C.c_long, C.c_int, dp]
        L.ref_rv2coe.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefSatellites:
    """An array of reference `perturb::Satellite` objects."""

    def __init__(self, n):
        self.n = int(n)
        self.h = lib().ref_alloc(self.n)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_free(self.h)
            self.h = None

    # -- construction routes -------------------------------------------------------
    @classmethod
    def from_tle_lines(cls, lines1, lines2, grav=1):
        """Satellite::from_tle (Vallado parser, opsmode 'i')."""
        self = cls(len(lines1))
        self.init_error = np.array(
            [lib().ref_init_from_tle(self.h, i, a.encode(), b.encode(), grav)
             for i, (a, b) in enumerate(zip(lines1, lines2))], dtype=np.int32)
        return self

    @classmethod
    def from_parsed(cls, lines1, lines2, grav=1):
        """TwoLineElement::parse + Satellite(TwoLineElement). Returns (sats, parse_errors)."""
        self = cls(len(lines1))
        perr = np.zeros(self.n, dtype=np.int32)
        ierr = np.zeros(self.n, dtype=np.int32)
        self.tles = []
        for i, (a, b) in enumerate(zip(lines1, lines2)):
            t = RefTle()
            perr[i] = lib().ref_parse_tle(a.encode(), b.encode(), C.byref(t))
            self.tles.append(t)
            ierr[i] = lib().ref_init_from_parsed(self.h, i, C.byref(t), grav)
        self.init_error = ierr
        return self, perr

    @classmethod
    def from_verif(cls, lines1, lines2, opsmode="a", grav=1):
        """twoline2rv('v','e',opsmode) as tests/test_perturb.cpp:28-50; keeps the grids."""
        self = cls(len(lines1))
        self.grids = []
        ierr = np.zeros(self.n, dtype=np.int32)
        for i, (a, b) in enumerate(zip(lines1, lines2)):
            s, e, d = C.c_double(), C.c_double(), C.c_double()
            ierr[i] = lib().ref_init_verif(
                self.h, i, a.encode(), b.encode(), opsmode.encode(), grav,
                C.byref(s), C.byref(e), C.byref(d))
            self.grids.append((s.value, e.value, d.value))
        self.init_error = ierr
        return self

    @classmethod
    def from_raw(cls, el, grav=1, opsmode="i"):
        """sgp4init on converted elements. `el` maps name -> array (see ref_init_raw)."""
        n = len(el["no_kozai"])
        self = cls(n)
        ierr = np.zeros(n, dtype=np.int32)
        for i in range(n):
            ierr[i] = lib().ref_init_raw(
                self.h, i, grav, opsmode.encode(),
                *[float(el[k][i]) for k in (
                    "jdsatepoch jdsatepochF bstar ndot nddot ecco argpo inclo mo no_kozai "
                    "nodeo").split()])
        self.init_error = ierr
        return self

    # -- queries -------------------------------------------------------------------
    def epoch(self):
        jd = np.zeros(self.n)
        fr = np.zeros(self.n)
        a, b = C.c_double(), C.c_double()
        for i in range(self.n):
            lib().ref_epoch(self.h, i, C.byref(a), C.byref(b))
            jd[i], fr[i] = a.value, b.value
        return jd, fr

    def fields(self):
        out = np.zeros((self.n, len(FIELD_NAMES)))
        for i in range(self.n):
            k = lib().ref_dump_fields(self.h, i, _dp(out[i]))
            assert k == len(FIELD_NAMES)
        return out

    def propagate_mins_one(self, i, mins):
        r = np.zeros(3)
        v = np.zeros(3)
        e = lib().ref_propagate_mins(self.h, i, float(mins), _dp(r), _dp(v))
        return e, r, v

    def propagate_jd_one(self, i, jd, frac):
        r = np.zeros(3)
        v = np.zeros(3)
        e = lib().ref_propagate_jd(self.h, i, float(jd), float(frac), _dp(r), _dp(v))
        return e, r, v

    def propagate_grid(self, ta, tb=None, use_jd=False, nthreads=1, want_state=True):
        """All satellites x all times. Returns (err[n,m], pos[n,m,3], vel[n,m,3], seconds)."""
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
        secs = lib().ref_propagate_grid(
            self.h, self.n, _dp(ta), _dp(tb), m, int(bool(use_jd)), int(nthreads), pp, vp,
            err.ctypes.data_as(C.POINTER(C.c_int)))
        return err, pos, vel, secs


def read_verif_file(path):
    """(line1, line2) pairs of SGP4-VER.TLE, comments skipped (test_perturb.cpp:343-347)."""
    l1, l2 = [], []
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    it = iter(lines)
    for a in it:
        if a.startswith("#"):
            continue
        b = next(it)
        l1.append(a)
        l2.append(b)
    return l1, l2


def rv2coe(r, v, grav=1):
    """ClassicalOrbitalElements(StateVector, grav) for states r, v [..., 3] -> [..., 11]."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(r.shape[:-1] + (11,))
    lib().ref_rv2coe(_dp(r), _dp(v), r.size // 3, grav, _dp(out))
    return out
