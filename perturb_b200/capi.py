"""ctypes view of the C ABI in include/perturb_gpu.h (the product, libperturb_gpu.so) and of the
measurement tooling in include/perturb_benchtools.h / perturb_synth.h (libperturb_benchtools.so:
synthetic catalogues, FP64 / D2H peak probes, the device-math test hook).

Both libraries are built in-tree (see csrc/Makefile). There is no CPU fallback: if the product
library is missing, loading fails loudly. `tools()` never maps the product library, so the
benchmark's reference arm can take the synthetic catalogue without touching it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PERTURB_GPU_LIB") or os.path.join(_HERE, "libperturb_gpu.so")
TOOLS_PATH = os.path.join(_HERE, "libperturb_benchtools.so")

OK = 0
WGS72_OLD, WGS72, WGS84 = 0, 1, 2
SYNTH_LEO, SYNTH_MIXED, SYNTH_DEEP, SYNTH_MEGA = 2, 3, 4, 5
SYNTH_LINE_STRIDE = 70

TLE_FIELDS = (
    "epoch_year epoch_day_of_year n_dot n_ddot b_star inclination raan eccentricity "
    "arg_of_perigee mean_anomaly mean_motion"
).split()


class Tle(C.Structure):
    """perturb_gpu_tle: numeric members of perturb::TwoLineElement (tle.hpp:65-91)."""
    _fields_ = [(n, C.c_double) for n in TLE_FIELDS]


# Every symbol include/perturb_gpu.h declares ...
EXPORTS = (
    "perturb_gpu_init perturb_gpu_init_from_text perturb_gpu_init_records perturb_gpu_free perturb_gpu_propagate "
    "perturb_gpu_propagate_from_epoch perturb_gpu_propagate_states perturb_gpu_propagate_elements perturb_gpu_propagate_summary perturb_gpu_rv2coe perturb_gpu_synchronize perturb_gpu_device_count perturb_gpu_host_alloc perturb_gpu_host_free perturb_gpu_device_alloc perturb_gpu_device_free perturb_gpu_copy perturb_gpu_size "
    "perturb_gpu_device perturb_gpu_last_error perturb_gpu_epoch perturb_gpu_orbit_class "
    "perturb_gpu_record_field_count perturb_gpu_record_field_name perturb_gpu_record_field "
    "perturb_gpu_last_launch_count perturb_gpu_strerror perturb_gpu_parse_tles perturb_gpu_parse_tle "
    "perturb_gpu_propagate_into perturb_gpu_elsetrec_size perturb_gpu_export_records "
    "perturb_gpu_write_back perturb_gpu_propagate_visibility perturb_gpu_propagate_proximity "
    "perturb_gpu_host_register perturb_gpu_host_unregister"
).split()
# ... and include/perturb_benchtools.h + include/perturb_synth.h (libperturb_benchtools.so)
TOOLS_EXPORTS = (
    "perturb_gpu_measure_fp64_peak perturb_gpu_measure_d2h perturb_gpu_measure_d2h_multi "
    "perturb_gpu_math_probe "
    "perturb_synth_catalog perturb_synth_time_grid"
).split()

_lib = None
_tools = None


class Summary(C.Structure):
    """perturb_gpu_summary"""
    _fields_ = [("r_min", C.c_double), ("r_max", C.c_double), ("k_min", C.c_int32),
                ("k_max", C.c_int32), ("n_ok", C.c_int32), ("first_error", C.c_int32),
                ("k_first_error", C.c_int32), ("reserved", C.c_int32)]


class PerturbGpuError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PerturbGpuError(
            "%s is missing: build it with `make -C perturb_b200/csrc` (or "
            "__graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.perturb_gpu_init.argtypes = [
        vp, C.c_size_t, C.c_int, C.c_char, C.c_int, C.POINTER(vp), ip]
    L.perturb_gpu_init_records.argtypes = L.perturb_gpu_init.argtypes
    L.perturb_gpu_init_from_text.argtypes = [
        vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_char, C.c_int, C.POINTER(vp), ip, ip]
    L.perturb_gpu_free.argtypes = [vp]
    L.perturb_gpu_free.restype = None
    L.perturb_gpu_propagate.argtypes = [
        vp, vp, vp, C.c_size_t, vp, vp, C.c_size_t, C.c_size_t, vp]
    L.perturb_gpu_propagate_from_epoch.argtypes = [
        vp, vp, C.c_size_t, vp, vp, C.c_size_t, C.c_size_t, vp]
    L.perturb_gpu_propagate_states.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, C.c_size_t, vp]
    L.perturb_gpu_propagate_elements.argtypes = [
        vp, vp, vp, C.c_size_t, vp, vp, vp, C.c_size_t, C.c_size_t, vp]
    L.perturb_gpu_propagate_summary.argtypes = [vp, vp, vp, C.c_size_t, vp, vp]
    L.perturb_gpu_rv2coe.argtypes = [
        vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp, C.c_size_t, vp]
    L.perturb_gpu_synchronize.argtypes = [vp]
    L.perturb_gpu_device_count.argtypes = [C.POINTER(C.c_int)]
    L.perturb_gpu_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.perturb_gpu_host_free.argtypes = [vp]
    L.perturb_gpu_host_free.restype = None
    L.perturb_gpu_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.perturb_gpu_device_free.argtypes = [C.c_int, vp]
    L.perturb_gpu_device_free.restype = None
    L.perturb_gpu_copy.argtypes = [vp, vp, C.c_size_t]
    L.perturb_gpu_size.argtypes = [vp]
    L.perturb_gpu_size.restype = C.c_size_t
    L.perturb_gpu_device.argtypes = [vp]
    L.perturb_gpu_last_error.argtypes = [vp, ip]
    L.perturb_gpu_epoch.argtypes = [vp, dp, dp]
    L.perturb_gpu_orbit_class.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.perturb_gpu_record_field_name.argtypes = [C.c_int]
    L.perturb_gpu_record_field_name.restype = C.c_char_p
    L.perturb_gpu_record_field.argtypes = [vp, C.c_int, dp]
    L.perturb_gpu_last_launch_count.argtypes = [vp]
    L.perturb_gpu_strerror.restype = C.c_char_p
    L.perturb_gpu_parse_tles.argtypes = [
        C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(Tle), ip]
    try:
        L.perturb_gpu_propagate_into.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, C.c_size_t]
        L.perturb_gpu_export_records.argtypes = [vp, vp, C.c_size_t]
        L.perturb_gpu_write_back.argtypes = [vp, C.c_double, C.c_double, C.c_int, vp, C.c_size_t]
        L.perturb_gpu_propagate_visibility.argtypes = [
            vp, vp, vp, C.c_size_t, vp, C.c_size_t, vp, vp, C.c_size_t, vp]
        L.perturb_gpu_propagate_proximity.argtypes = [
            vp, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_double, vp, vp]
        L.perturb_gpu_host_register.argtypes = [vp, C.c_size_t]
        L.perturb_gpu_host_unregister.argtypes = [vp]
        L.perturb_gpu_host_unregister.restype = None
    except AttributeError:
        # an older comparison build loaded through PERTURB_GPU_LIB (kernel sweeps only)
        if "PERTURB_GPU_LIB" not in os.environ:
            raise
    _lib = L
    return L


def tools():
    """libperturb_benchtools.so: workload generator and measurement probes (not the product)."""
    global _tools
    if _tools is not None:
        return _tools
    if not os.path.exists(TOOLS_PATH):
        raise PerturbGpuError("%s is missing: build it with `make -C perturb_b200/csrc`" % TOOLS_PATH)
    T = C.CDLL(TOOLS_PATH)
    dp = C.POINTER(C.c_double)
    T.perturb_gpu_measure_fp64_peak.argtypes = [C.c_int, dp]
    T.perturb_gpu_measure_d2h.argtypes = [C.c_int, C.c_size_t, C.c_int, dp, dp]
    T.perturb_gpu_measure_d2h_multi.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_uint, dp]
    T.perturb_gpu_math_probe.argtypes = [C.c_int, dp, dp, dp, C.c_size_t]
    T.perturb_synth_catalog.argtypes = [
        C.c_int, C.c_size_t, C.c_uint64, C.c_char_p, C.c_char_p]
    T.perturb_synth_time_grid.argtypes = [C.c_size_t, dp, dp]
    _tools = T
    return T


def check(status, what):
    if status != OK:
        msg = lib().perturb_gpu_strerror()
        raise PerturbGpuError("%s failed (%d): %s" % (what, status, (msg or b"").decode()))
