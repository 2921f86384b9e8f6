"""Python mirror of `perturb::SatelliteBatch` (include/perturb/satellite_batch.hpp).

Same names and argument meaning as the reference's `Satellite` interface for this path
(include/perturb/perturb.hpp:241-322), batched:

    Satellite::from_tle(line_1, line_2, grav)     -> SatelliteBatch.from_tles(lines1, lines2, grav)
    Satellite(const TwoLineElement&, grav)        -> SatelliteBatch.from_parsed(tles, grav)
    Satellite(sgp4::elsetrec)                     -> SatelliteBatch.from_records(records, grav)
    sat.last_error()                              -> batch.last_error()          [n]
    sat.epoch()                                   -> batch.epoch()               ([n], [n])
    sat.propagate(JulianDate, sv)                 -> batch.propagate(jd, jd_frac)
    sat.propagate_from_epoch(mins, sv)            -> batch.propagate_from_epoch(mins)

Everything goes through the C ABI (perturb_b200/capi.py); torch is used only to own device /
pinned buffers. Never raises on per-satellite problems: those are Sgp4Error codes in the
returned arrays, like the reference.
"""
import ctypes as C

import numpy as np

from . import capi


def _as_line_block(lines, stride=capi.SYNTH_LINE_STRIDE):
    """list[str] -> one bytes block of fixed-stride records (NUL padded)."""
    buf = bytearray(len(lines) * stride)
    for i, ln in enumerate(lines):
        b = ln.encode("ascii", "replace")[: stride - 1]
        buf[i * stride: i * stride + len(b)] = b
    return bytes(buf)


def parse_tles(lines1, lines2, flavor=1):
    """Host TLE text front end. Returns (ctypes Tle array, status[n]).

    flavor 1 = numeric conventions of `Satellite::from_tle` / twoline2rv,
    flavor 0 = those of `TwoLineElement::parse`."""
    n = len(lines1)
    if isinstance(lines1, (bytes, bytearray)):
        raise TypeError("pass lists of strings, or use parse_tle_blocks for raw blocks")
    return parse_tle_blocks(_as_line_block(lines1), _as_line_block(lines2), n, flavor)


def parse_tle_blocks(block1, block2, n, flavor=1, stride=capi.SYNTH_LINE_STRIDE):
    out = (capi.Tle * n)()
    status = np.zeros(n, dtype=np.int32)
    st = capi.lib().perturb_gpu_parse_tles(
        block1, block2, stride, n, flavor, out, status.ctypes.data_as(C.POINTER(C.c_int32)))
    capi.check(st, "perturb_gpu_parse_tles")
    return out, status


def synth_catalog(config, n, seed=0):
    """Synthetic catalogue of BASELINE config `config` as two fixed-stride line blocks."""
    b1 = C.create_string_buffer(n * capi.SYNTH_LINE_STRIDE)
    b2 = C.create_string_buffer(n * capi.SYNTH_LINE_STRIDE)
    st = capi.tools().perturb_synth_catalog(config, n, seed, b1, b2)
    if st != 0:
        raise ValueError("bad synthetic catalogue arguments")
    return b1.raw, b2.raw


def block_to_lines(block, n, stride=capi.SYNTH_LINE_STRIDE):
    return [block[i * stride: i * stride + 69].decode("ascii") for i in range(n)]


def synth_time_grid(m):
    jd = np.zeros(m)
    fr = np.zeros(m)
    dp = C.POINTER(C.c_double)
    capi.tools().perturb_synth_time_grid(m, jd.ctypes.data_as(dp), fr.ctypes.data_as(dp))
    return jd, fr


def _ptr(x):
    """Raw address of a numpy array / torch tensor / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


def _dtype_name(x):
    return str(x.dtype).replace("torch.", "")


def _times(x, what):
    """A time-grid argument as the C ABI reads it: float64, one-dimensional, contiguous. Numpy
    arrays and Python sequences are converted; tensors must already be right (a silent copy of
    a device tensor would hide a caller's mistake)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray) or not hasattr(x, "data_ptr"):
        arr = np.ascontiguousarray(x, dtype=np.float64)
        if arr.ndim != 1:
            raise ValueError("%s must be one-dimensional" % what)
        return arr
    if _dtype_name(x) != "float64":
        raise TypeError("%s must be float64, got %s" % (what, _dtype_name(x)))
    if x.dim() != 1 or (x.numel() > 1 and x.stride(0) != 1):
        raise ValueError("%s must be a contiguous one-dimensional tensor" % what)
    return x


def _elem_strides(x):
    if isinstance(x, np.ndarray):
        return [st // x.itemsize for st in x.strides]
    return list(x.stride())


def _check_output(x, what, dtype, shape_min, last_contiguous=True):
    """An output buffer as the C ABI writes it: the right element type, at least the given
    shape, the last axis contiguous."""
    if _dtype_name(x) != dtype:
        raise TypeError("%s must be %s, got %s" % (what, dtype, _dtype_name(x)))
    shape = tuple(x.shape)
    if len(shape) != len(shape_min) or any(a < b for a, b in zip(shape, shape_min)):
        raise ValueError("%s has shape %s, need at least %s" % (what, shape, tuple(shape_min)))
    if 0 in shape:
        return  # nothing will be written; the strides of an empty array mean nothing
    if last_contiguous and shape[-1] > 1 and _elem_strides(x)[-1] != 1:
        raise ValueError("%s: the last axis must be contiguous" % what)


COE_NAMES = (
    "semilatus_rectum semimajor_axis eccentricity inclination raan arg_of_perigee "
    "true_anomaly mean_anomaly arg_of_latitude true_longitude longitude_of_periapsis").split()


def classical_elements(posvel, grav=capi.WGS72):
    """Batched `ClassicalOrbitalElements(sv, grav)`: posvel [6, n, m] (numpy or torch CUDA)
    -> coe [11, n, m] torch CUDA tensor, planes in COE_NAMES order."""
    import torch
    pv = torch.as_tensor(posvel, dtype=torch.float64)
    if not pv.is_cuda:
        pv = pv.cuda()
    if pv.stride(2) != 1:
        pv = pv.contiguous()
    _, n, m = pv.shape
    row, plane = pv.stride(1), pv.stride(0)
    if n <= 1:
        row = max(row, m)
    plane = max(plane, n * row)
    coe = torch.empty((11, n, row), dtype=torch.float64, device=pv.device)
    st = capi.lib().perturb_gpu_rv2coe(pv.data_ptr(), n, m, row, plane, int(grav),
                                       coe.data_ptr(), n * row, None)
    capi.check(st, "perturb_gpu_rv2coe")
    torch.cuda.synchronize(pv.device)
    return coe[:, :, :m]


class SatelliteBatch:
    def __init__(self, handle, n):
        self._h = handle
        self.n = n

    # ---- construction -----------------------------------------------------------------
    @classmethod
    def _create(cls, fn, src, n, grav, opsmode, device):
        h = C.c_void_p()
        init_error = np.zeros(n, dtype=np.int32)
        st = fn(src, n, int(grav), opsmode.encode(), int(device), C.byref(h),
                init_error.ctypes.data_as(C.POINTER(C.c_int32)))
        capi.check(st, fn.__name__)
        self = cls(h, n)
        self._init_error = init_error
        return self

    @classmethod
    def from_parsed(cls, tles, grav=capi.WGS72, opsmode="i", device=0):
        """`tles`: ctypes array of capi.Tle (TLE units), or float64 array [n, 11]."""
        if isinstance(tles, np.ndarray):
            arr = np.ascontiguousarray(tles, dtype=np.float64)
            assert arr.ndim == 2 and arr.shape[1] == len(capi.TLE_FIELDS)
            n, src = arr.shape[0], arr.ctypes.data
            self = cls._create(capi.lib().perturb_gpu_init, src, n, grav, opsmode, device)
            return self
        n = len(tles)
        return cls._create(capi.lib().perturb_gpu_init, C.addressof(tles) if n else None, n,
                           grav, opsmode, device)

    @classmethod
    def from_tles(cls, lines1, lines2, grav=capi.WGS72, opsmode="i", device=0, flavor=1):
        tles, status = parse_tles(lines1, lines2, flavor)
        self = cls.from_parsed(tles, grav, opsmode, device)
        self.parse_status = status
        return self

    @classmethod
    def from_text(cls, lines1, lines2, grav=capi.WGS72, opsmode="i", device=0, flavor=1, n=None):
        """Like from_tles, but the TLE text is parsed on the device (kernel) instead of the
        host. lines1/lines2: lists of strings, or fixed-stride byte blocks with `n` given."""
        if isinstance(lines1, (bytes, bytearray)):
            b1, b2 = bytes(lines1), bytes(lines2)
            if n is None or min(len(b1), len(b2)) < n * capi.SYNTH_LINE_STRIDE:
                raise ValueError("line blocks shorter than n * %d bytes" % capi.SYNTH_LINE_STRIDE)
        else:
            n = len(lines1)
            b1, b2 = _as_line_block(lines1), _as_line_block(lines2)
        h = C.c_void_p()
        init_error = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        st = capi.lib().perturb_gpu_init_from_text(
            b1, b2, capi.SYNTH_LINE_STRIDE, n, int(flavor), int(grav), opsmode.encode(),
            int(device), C.byref(h), init_error.ctypes.data_as(ip), status.ctypes.data_as(ip))
        capi.check(st, "perturb_gpu_init_from_text")
        self = cls(h, n)
        self._init_error = init_error
        self.parse_status = status
        return self

    @classmethod
    def from_records(cls, records, grav=capi.WGS72, opsmode="i", device=0):
        """records: float64 [record_field_count, n], field order = record_field_names()."""
        arr = np.ascontiguousarray(records, dtype=np.float64)
        assert arr.ndim == 2 and arr.shape[0] == capi.lib().perturb_gpu_record_field_count()
        return cls._create(capi.lib().perturb_gpu_init_records, arr.ctypes.data, arr.shape[1],
                           grav, opsmode, device)

    def close(self):
        if self._h:
            capi.lib().perturb_gpu_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- queries ----------------------------------------------------------------------
    def last_error(self):
        out = np.zeros(self.n, dtype=np.int32)
        capi.check(capi.lib().perturb_gpu_last_error(
            self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), "perturb_gpu_last_error")
        return out

    def epoch(self):
        jd = np.zeros(self.n)
        fr = np.zeros(self.n)
        dp = C.POINTER(C.c_double)
        capi.check(capi.lib().perturb_gpu_epoch(
            self._h, jd.ctypes.data_as(dp), fr.ctypes.data_as(dp)), "perturb_gpu_epoch")
        return jd, fr

    def orbit_class(self):
        out = np.zeros(self.n, dtype=np.uint8)
        capi.check(capi.lib().perturb_gpu_orbit_class(
            self._h, out.ctypes.data_as(C.POINTER(C.c_uint8))), "perturb_gpu_orbit_class")
        return out

    @staticmethod
    def record_field_names():
        L = capi.lib()
        return [L.perturb_gpu_record_field_name(i).decode()
                for i in range(L.perturb_gpu_record_field_count())]

    def record_field(self, name):
        idx = self.record_field_names().index(name)
        out = np.zeros(self.n)
        capi.check(capi.lib().perturb_gpu_record_field(
            self._h, idx, out.ctypes.data_as(C.POINTER(C.c_double))), "perturb_gpu_record_field")
        return out

    def records(self):
        return np.stack([self.record_field(nm) for nm in self.record_field_names()])

    def last_launch_count(self):
        return capi.lib().perturb_gpu_last_launch_count(self._h)

    def synchronize(self):
        capi.check(capi.lib().perturb_gpu_synchronize(self._h), "perturb_gpu_synchronize")

    # ---- the hot path -----------------------------------------------------------------
    def _outputs(self, m, posvel, err):
        if posvel is None:
            posvel = np.empty((6, self.n, m), dtype=np.float64)
        if err is None:
            err = np.empty((self.n, m), dtype=np.int32)
        return posvel, err

    @staticmethod
    def _strides(posvel, n, m):
        if isinstance(posvel, np.ndarray):
            es = posvel.itemsize
            s = [x // es for x in posvel.strides]
        else:
            s = list(posvel.stride())
        if n == 0 or m == 0:
            return max(m, 1), max(n * m, 1)
        assert s[2] == 1 or m == 1, "time must be the contiguous axis"
        row = s[1] if n > 1 else max(m, s[1] if s[1] >= m else m)
        plane = s[0]
        return row, max(plane, n * row)  # row_stride, plane_stride (doubles)

    def _checked_planes(self, m, posvel, err):
        """Validated (posvel, err, row_stride, plane_stride) of a planes-mode call."""
        posvel, err = self._outputs(m, posvel, err)
        _check_output(posvel, "posvel", "float64", (6, self.n, m))
        _check_output(err, "err", "int32", (self.n, m))
        row, plane = self._strides(posvel, self.n, m)
        if self.n > 1 and m > 0 and _elem_strides(err)[0] != row:
            raise ValueError("err row stride (%d) differs from posvel row stride (%d)"
                             % (_elem_strides(err)[0], row))
        return posvel, err, row, plane

    def propagate(self, jd, jd_frac, posvel=None, err=None, stream=None, sync=True):
        """All satellites x all JulianDate(jd[k], jd_frac[k]).

        Returns (posvel [6, n, m] = x,y,z [km], vx,vy,vz [km/s]; err [n, m] Sgp4Error).
        jd/jd_frac/posvel/err may be numpy arrays (host) or torch CUDA tensors (device)."""
        jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
        m = int(jd.shape[0])
        if int(jd_frac.shape[0]) != m:
            raise ValueError("jd and jd_frac differ in length")
        posvel, err, row, plane = self._checked_planes(m, posvel, err)
        st = capi.lib().perturb_gpu_propagate(
            self._h, _ptr(jd), _ptr(jd_frac), m, _ptr(posvel), _ptr(err), row, plane, stream)
        capi.check(st, "perturb_gpu_propagate")
        if sync:
            self.synchronize()
        return posvel, err

    # perturb::StateVector (perturb.hpp:198-205): JulianDate epoch; Vec3 position; Vec3 velocity
    STATE_DTYPE = np.dtype([("epoch_jd", "f8"), ("epoch_jd_frac", "f8"), ("position", "f8", 3),
                            ("velocity", "f8", 3)])

    def propagate_states(self, jd, jd_frac=None, states=None, err=None, stream=None, sync=True):
        """propagate (or propagate_from_epoch when jd_frac is None, jd = minutes) into an array
        of 64-byte perturb::StateVector records: returns (states [n, m] STATE_DTYPE, err [n, m]).
        `states` / `err` may be given: numpy arrays (host, STATE_DTYPE / int32) or CUDA torch
        tensors (float64 [n, m, 8] / int32 [n, m])."""
        jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
        m = int(jd.shape[0])
        if states is None:
            states = np.empty((self.n, m), dtype=self.STATE_DTYPE)
            if err is None:
                err = np.empty((self.n, m), dtype=np.int32)
        if err is None:
            raise ValueError("give err together with states")
        _check_output(err, "err", "int32", (self.n, m))
        if hasattr(states, "is_cuda"):
            _check_output(states, "states", "float64", (self.n, m, 8))
            row = int(states.stride(0) // 8)
        else:
            if states.dtype != self.STATE_DTYPE:
                raise TypeError("states must have STATE_DTYPE")
            _check_output(states, "states", str(states.dtype), (self.n, m))
            row = int(states.strides[0] // 64)
        if self.n > 1 and m > 0 and _elem_strides(err)[0] != row:
            raise ValueError("err row stride differs from the states row stride")
        st = capi.lib().perturb_gpu_propagate_states(
            self._h, _ptr(jd), _ptr(jd_frac), m, _ptr(states), _ptr(err), row, stream)
        capi.check(st, "perturb_gpu_propagate_states")
        if sync:
            self.synchronize()
        return states, err

    def propagate_elements(self, jd, jd_frac=None, stream=None, host=False):
        """propagate (or propagate_from_epoch when jd_frac is None, jd = minutes) that also
        returns the mean elements sgp4() leaves in sat_rec: (posvel [6,n,m], err [n,m],
        elements [7,n,m] = am, em, im, Om, om, mm, nm), all as torch CUDA tensors -- or, with
        host=True, as numpy arrays filled through the C ABI's host-buffer path."""
        if host:
            jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
            m = int(jd.shape[0])
            posvel = np.empty((6, self.n, m))
            elements = np.empty((7, self.n, m))
            err = np.empty((self.n, m), dtype=np.int32)
            st = capi.lib().perturb_gpu_propagate_elements(
                self._h, _ptr(jd), _ptr(jd_frac), m, _ptr(posvel), _ptr(err), _ptr(elements),
                max(m, 1), max(self.n * m, 1), stream)
            capi.check(st, "perturb_gpu_propagate_elements")
            return posvel, err, elements
        import torch
        dev = torch.device("cuda", capi.lib().perturb_gpu_device(self._h))
        m = int(jd.shape[0])
        d_a = torch.as_tensor(jd, dtype=torch.float64, device=dev)
        d_b = None if jd_frac is None else torch.as_tensor(jd_frac, dtype=torch.float64, device=dev)
        m2 = m + (m & 1)
        posvel = torch.empty((6, self.n, m2), dtype=torch.float64, device=dev)[:, :, :m]
        elements = torch.empty((7, self.n, m2), dtype=torch.float64, device=dev)[:, :, :m]
        err = torch.empty((self.n, m2), dtype=torch.int32, device=dev)[:, :m]
        st = capi.lib().perturb_gpu_propagate_elements(
            self._h, _ptr(d_a), _ptr(d_b), m, _ptr(posvel), _ptr(err), _ptr(elements), m2,
            self.n * m2, stream)
        capi.check(st, "perturb_gpu_propagate_elements")
        self.synchronize()
        return posvel, err, elements

    SUMMARY_DTYPE = np.dtype([("r_min", "f8"), ("r_max", "f8"), ("k_min", "i4"), ("k_max", "i4"),
                              ("n_ok", "i4"), ("first_error", "i4"), ("k_first_error", "i4"),
                              ("reserved", "i4")])

    def propagate_summary(self, jd, jd_frac=None, stream=None):
        """Fused consumer: per-satellite envelope of a propagation (numpy structured array [n],
        SUMMARY_DTYPE) without materialising the [n, m] states. jd_frac None: jd = minutes."""
        jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
        out = np.zeros(self.n, dtype=self.SUMMARY_DTYPE)
        assert out.dtype.itemsize == C.sizeof(capi.Summary)
        st = capi.lib().perturb_gpu_propagate_summary(
            self._h, _ptr(jd), _ptr(jd_frac), int(jd.shape[0]), out.ctypes.data, stream)
        capi.check(st, "perturb_gpu_propagate_summary")
        self.synchronize()
        return out

    def propagate_summary_device(self, jd, jd_frac, out, stream=None):
        """Same with a device-resident result: `out` is a CUDA torch.uint8 tensor of
        n * SUMMARY_DTYPE.itemsize bytes; asynchronous on `stream`."""
        assert out.is_cuda and out.numel() * out.element_size() >= self.n * self.SUMMARY_DTYPE.itemsize
        st = capi.lib().perturb_gpu_propagate_summary(
            self._h, _ptr(jd), _ptr(jd_frac), int(jd.shape[0]), out.data_ptr(), stream)
        capi.check(st, "perturb_gpu_propagate_summary")

    STATION_DTYPE = np.dtype([("latitude_rad", "f8"), ("longitude_rad", "f8"), ("altitude_km", "f8"),
                              ("min_elevation_rad", "f8")])
    PASS_DTYPE = np.dtype([("max_sin_elevation", "f8"), ("k_max_elevation", "i4"), ("n_visible", "i4"),
                           ("n_passes", "i4"), ("k_first_rise", "i4"), ("k_last_visible", "i4"),
                           ("reserved", "i4")])
    APPROACH_DTYPE = np.dtype([("d_min_km", "f8"), ("k_min", "i4"), ("n_below", "i4")])

    def propagate_visibility(self, jd, jd_frac, stations, max_passes=0, stream=None):
        """Fused ground-station visibility screen (perturb_gpu_propagate_visibility): `stations`
        is a sequence of (latitude_rad, longitude_rad, altitude_km, min_elevation_rad). Returns
        (summary [n, S] PASS_DTYPE, passes [n, S, max_passes, 2] int32 or None) without
        materialising the [n, m] states."""
        jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
        st_arr = np.zeros(len(stations), dtype=self.STATION_DTYPE)
        for j, row in enumerate(stations):
            st_arr[j] = tuple(float(x) for x in row)
        out = np.zeros((self.n, len(stations)), dtype=self.PASS_DTYPE)
        passes = np.zeros((self.n, len(stations), max_passes, 2), dtype=np.int32) if max_passes else None
        st = capi.lib().perturb_gpu_propagate_visibility(
            self._h, _ptr(jd), _ptr(jd_frac), int(jd.shape[0]), st_arr.ctypes.data, len(stations),
            out.ctypes.data, _ptr(passes), int(max_passes), stream)
        capi.check(st, "perturb_gpu_propagate_visibility")
        self.synchronize()
        return out, passes

    def propagate_proximity(self, jd, jd_frac, primaries, threshold_km, stream=None):
        """Fused closest-approach screen (perturb_gpu_propagate_proximity): `primaries` [P, 3, m]
        float64 TEME positions on the same grid (numpy, or a CUDA tensor on the batch's device).
        Returns [n, P] APPROACH_DTYPE. jd_frac None: jd = minutes."""
        jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
        m = int(jd.shape[0])
        if isinstance(primaries, np.ndarray):
            primaries = np.ascontiguousarray(primaries, dtype=np.float64)
        _check_output(primaries, "primaries", "float64", (1, 3, m))
        if tuple(primaries.shape[1:]) != (3, m) or _elem_strides(primaries) != [3 * m, m, 1]:
            raise ValueError("primaries must be a contiguous [P, 3, m] array")
        n_prim = int(primaries.shape[0])
        out = np.zeros((self.n, n_prim), dtype=self.APPROACH_DTYPE)
        st = capi.lib().perturb_gpu_propagate_proximity(
            self._h, _ptr(jd), _ptr(jd_frac), m, _ptr(primaries), n_prim, float(threshold_km),
            out.ctypes.data, stream)
        capi.check(st, "perturb_gpu_propagate_proximity")
        self.synchronize()
        return out

    def propagate_into(self, jd, jd_frac, states, err=None):
        """The drop-in loop `errs[i, k] = sats[i].propagate(times[k], svs[i, k])` into the
        caller's own HOST arrays (perturb_gpu_propagate_into): `states` [n, m] STATE_DTYPE is
        updated in place -- epoch always, position / velocity for codes NONE and DECAYED only,
        otherwise left as they were, like the reference. jd_frac None: jd = minutes."""
        jd, jd_frac = _times(jd, "jd"), _times(jd_frac, "jd_frac")
        m = int(jd.shape[0])
        if not isinstance(states, np.ndarray) or states.dtype != self.STATE_DTYPE:
            raise TypeError("states must be a numpy array of STATE_DTYPE")
        _check_output(states, "states", str(states.dtype), (self.n, m))
        row = int(states.strides[0] // 64)
        if err is not None:
            _check_output(err, "err", "int32", (self.n, m))
            if self.n > 1 and m > 0 and _elem_strides(err)[0] != row:
                raise ValueError("err row stride differs from the states row stride")
        st = capi.lib().perturb_gpu_propagate_into(
            self._h, _ptr(jd), _ptr(jd_frac), m, _ptr(states), _ptr(err), row)
        capi.check(st, "perturb_gpu_propagate_into")
        return states, err

    # ---- write-back into the reference's own records (N2) -------------------------------
    def export_records(self, records=None, stride=None):
        """The initialised `sgp4::elsetrec` of every satellite as the reference lays it out
        (perturb_gpu_export_records): uint8 array [n, stride], stride >= 1000."""
        size = capi.lib().perturb_gpu_elsetrec_size()
        if records is None:
            records = np.zeros((self.n, stride or size), dtype=np.uint8)
        _check_output(records, "records", "uint8", (self.n, size))
        st = capi.lib().perturb_gpu_export_records(self._h, records.ctypes.data, records.strides[0])
        capi.check(st, "perturb_gpu_export_records")
        return records

    def write_back(self, records, jd, jd_frac=None):
        """What ONE propagate call at JulianDate(jd, jd_frac) (or minutes `jd` when jd_frac is
        None) leaves in `sat_rec`, written into `records` [n, stride] uint8 in place."""
        size = capi.lib().perturb_gpu_elsetrec_size()
        _check_output(records, "records", "uint8", (self.n, size))
        st = capi.lib().perturb_gpu_write_back(
            self._h, float(jd), 0.0 if jd_frac is None else float(jd_frac),
            0 if jd_frac is None else 1, records.ctypes.data, records.strides[0])
        capi.check(st, "perturb_gpu_write_back")
        return records

    def propagate_from_epoch(self, mins, posvel=None, err=None, stream=None, sync=True):
        mins = _times(mins, "mins")
        m = int(mins.shape[0])
        posvel, err, row, plane = self._checked_planes(m, posvel, err)
        st = capi.lib().perturb_gpu_propagate_from_epoch(
            self._h, _ptr(mins), m, _ptr(posvel), _ptr(err), row, plane, stream)
        capi.check(st, "perturb_gpu_propagate_from_epoch")
        if sync:
            self.synchronize()
        return posvel, err
