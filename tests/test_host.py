"""CPU-side tests of the product's host logic and of the C-ABI surface (no compute calls)."""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import perturb_b200 as pb
from helpers import load_synth_golden, load_ver_golden, ver_lines
from perturb_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = load_ver_golden()


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(perturb_(?:gpu|synth)_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _declared_functions("perturb_gpu.h")
    assert len(declared) >= 19
    lib = C.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libperturb_gpu.so does not export " + name
    assert sorted(declared) == sorted(capi.EXPORTS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH]).decode()
    exported = set(re.findall(r" T (perturb_\w+)", out))
    assert set(declared) <= exported
    # the product carries no bench / test hooks: those live in libperturb_benchtools.so
    assert not (exported & set(capi.TOOLS_EXPORTS)), exported & set(capi.TOOLS_EXPORTS)


def test_benchtools_library_exports_its_header_and_is_not_the_product():
    declared = sorted(set(_declared_functions("perturb_benchtools.h") + _declared_functions("perturb_synth.h")))
    assert declared == sorted(capi.TOOLS_EXPORTS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.TOOLS_PATH]).decode()
    exported = set(re.findall(r" T (perturb_\w+)", out))
    assert set(declared) <= exported
    assert not (exported & set(capi.EXPORTS))
    needed = subprocess.check_output(["readelf", "-d", capi.TOOLS_PATH]).decode()
    assert "libperturb_gpu" not in needed


def test_library_is_sm100a_and_has_both_kernels():
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH]).decode()
    assert "sm_100a" in out
    sym = subprocess.check_output(["cuobjdump", "-symbols", capi.LIB_PATH]).decode()
    assert "sgp4init_kernel" in sym and "sgp4_prop_kernel" in sym


def test_record_field_table_matches_reference_record():
    L = capi.lib()
    n = L.perturb_gpu_record_field_count()
    names = [L.perturb_gpu_record_field_name(i).decode() for i in range(n)]
    assert names == [str(x) for x in G["field_names"]]
    assert L.perturb_gpu_record_field_name(n) is None


def test_product_never_touches_the_oracle():
    """No CPU fallback: nothing under perturb_b200/ or include/ may reference oracle/."""
    bad = []
    for base in ("perturb_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    text = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(^|\W)(import oracle|from oracle|oracle/|sgp4_oracle|liboracle)",
                                 text.replace("oracle/ref_driver.cpp:ref_dump_fields", "")):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_device_fails_loudly_not_silently():
    tles, _ = pb.parse_tles(*ver_lines(G))
    with pytest.raises(pb.PerturbGpuError) as ei:
        pb.SatelliteBatch.from_parsed(tles)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_argument_validation_without_gpu():
    L = capi.lib()
    h = C.c_void_p()
    assert L.perturb_gpu_init(None, 4, 1, b"i", 0, None, None) == -1
    assert L.perturb_gpu_init(None, 0, 1, b"x", 0, C.byref(h), None) == -1
    assert L.perturb_gpu_propagate(None, None, None, 1, None, None, 1, 1, None) == -1
    assert L.perturb_gpu_size(None) == 0
    assert L.perturb_gpu_parse_tles(None, None, 70, 1, 0, None, None) == -1
    assert L.perturb_gpu_parse_tles(b"x" * 70, b"y" * 70, 10, 1, 0, (capi.Tle * 1)(), None) == -1
    L.perturb_gpu_parse_tle.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(capi.Tle), C.c_void_p]
    assert L.perturb_gpu_parse_tle(None, b"y" * 70, (capi.Tle * 1)(), None) == -1
    assert L.perturb_gpu_parse_tle(b"short", b"y" * 70, (capi.Tle * 1)(), None) == 2  # INVALID_FORMAT
    assert L.perturb_gpu_copy(None, None, 8) == -1 and L.perturb_gpu_copy(None, None, 0) == 0
    p = C.c_void_p(1)
    assert L.perturb_gpu_device_alloc(0, 0, C.byref(p)) == 0 and p.value is None
    assert L.perturb_gpu_device_alloc(0, 64, None) == -1
    L.perturb_gpu_device_free(0, None)  # a no-op, like free(NULL)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_device_memory_helpers_fail_loudly_without_a_gpu():
    p = C.c_void_p()
    st = capi.lib().perturb_gpu_device_alloc(0, 1 << 20, C.byref(p))
    assert st in (-3, -4) and p.value is None  # NODEV / ALLOC, never a silent host buffer


# ---- TLE front end ---------------------------------------------------------------------
def test_parser_flavor0_equals_reference_twolineelement_parse():
    l1, l2 = ver_lines(G)
    tles, status = pb.parse_tles(l1, l2, flavor=0)
    ref_err = G["tle_parse_error"]
    for i in np.nonzero(ref_err == 0)[0]:
        assert status[i] == 0
        got = np.array([getattr(tles[i], k) for k in capi.TLE_FIELDS])
        assert np.array_equal(got, G["tle_parse_values"][i]), i
    # every status agrees, including the STR#3 card with a blank designator (88888) that the
    # reference's sscanf mis-tokenises and rejects with INVALID_VALUE
    assert np.array_equal(status, ref_err)
    assert set(status[ref_err == 4]) == {4} and 3 in set(status)


def test_parser_flavor1_equals_vallado_conversion():
    """twoline2rv's numbers, recovered from the reference's records (units undone exactly
    is not possible, so compare the converted elements the way sgp4init receives them)."""
    l1, l2 = ver_lines(G)
    tles, status = pb.parse_tles(l1, l2, flavor=1)
    assert (status == 0).all()
    names = [str(x) for x in G["field_names"]]
    F = G["pub_fields"]
    deg = np.pi / 180.0
    xp = 1440.0 / (2.0 * np.pi)
    for i, t in enumerate(tles):
        want = {"bstar": t.b_star, "ecco": t.eccentricity, "inclo": t.inclination * deg,
                "nodeo": t.raan * deg, "argpo": t.arg_of_perigee * deg,
                "mo": t.mean_anomaly * deg, "no_kozai": t.mean_motion / xp,
                "ndot": t.n_dot / (xp * 1440.0), "nddot": t.n_ddot / (xp * 1440.0 * 1440)}
        for k, v in want.items():
            assert F[i, names.index(k)] == v, (i, k)


def test_parser_rejects_short_and_bad_lines():
    l1, l2 = ver_lines(G)
    tl, st = pb.parse_tles([l1[0][:40], l1[1]], [l2[0], l2[1]], flavor=1)
    assert st[0] == 2 and np.isnan(tl[0].mean_motion) and st[1] == 0
    bad = l1[2][:68] + ("0" if l1[2][68] != "0" else "1")
    tl, st = pb.parse_tles([bad], [l2[2]], flavor=0)
    assert st[0] == 4 and np.isnan(tl[0].epoch_year)  # CHECKSUM_MISMATCH
    nospace = l1[2][:8] + "X" + l1[2][9:]
    _, st = pb.parse_tles([nospace], [l2[2]], flavor=0)
    assert st[0] == 1  # SHOULD_BE_SPACE comes first (tle.hpp:39-44)
    tl, st = pb.parse_tles([], [], flavor=0)
    assert len(tl) == 0 and st.size == 0


# ---- synthetic catalogues --------------------------------------------------------------
def test_synth_catalog_is_deterministic_and_matches_fixture_text():
    S = load_synth_golden()
    for cfg, n in ((2, 128), (3, 256), (4, 128), (5, 64)):
        b1, b2 = pb.synth_catalog(cfg, n)
        assert pb.block_to_lines(b1, n) == [str(x) for x in S["lines1_c%d" % cfg]]
        assert pb.block_to_lines(b2, n) == [str(x) for x in S["lines2_c%d" % cfg]]
        c1, c2 = pb.synth_catalog(cfg, n)
        assert (b1, b2) == (c1, c2)
        d1, _ = pb.synth_catalog(cfg, n, seed=7)
        assert d1 != b1


def test_synth_catalog_lines_are_valid_tles():
    for cfg in (2, 3, 4, 5):
        n = 2000
        b1, b2 = pb.synth_catalog(cfg, n)
        tles, status = pb.parse_tle_blocks(b1, b2, n, flavor=0)  # strict reader + checksums
        assert (status == 0).all(), (cfg, np.bincount(status))
        nmo = np.array([t.mean_motion for t in tles])
        ecc = np.array([t.eccentricity for t in tles])
        assert (nmo > 0.5).all() and (ecc >= 0).all() and (ecc < 1).all()
        if cfg == 4:
            assert ((nmo < 1.02).sum(), (ecc > 0.5).sum()) == (n // 2, n // 2)


def test_synth_full_size_fingerprint():
    """The C2 catalogue of BASELINE.json (30k) is one fixed byte string."""
    b1, b2 = pb.synth_catalog(2, 30000)
    h = hashlib.sha256(b1 + b2).hexdigest()
    again = hashlib.sha256(b"".join(pb.synth_catalog(2, 30000))).hexdigest()
    assert h == again
    jd, fr = pb.synth_time_grid(1440)
    assert jd[0] == 2461317.5 and fr[1] == 1.0 / 1440.0 and fr[-1] == 1439.0 / 1440.0


def test_abi_headers_are_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: both headers compile as C99 with -Wpedantic -Werror."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "perturb_gpu.h"\n#include "perturb_synth.h"\n#include "perturb_benchtools.h"\n'
                   'int main(void) { return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-fsyntax-only",
                           "-I" + os.path.join(ROOT, "include"), str(src)])


def test_reference_arm_never_maps_the_product_library():
    """bench.py --impl reference times the reference on the CPU: its process may map the
    benchtools library (catalogue text) and oracle/, never libperturb_gpu.so."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import bench\n"
        "r = bench.cpu_arm('c2', 1, 0, cores=2, budget_s=0.2)\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libperturb_benchtools.so' in maps\n"
        "assert 'libperturb_gpu.so' not in maps, 'product library mapped by the reference arm'\n"
        "print(r['kind'], r['value'] > 0)\n" % ROOT)
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    assert out.stdout.split()[-1] == "True"


def test_both_bench_arms_describe_the_same_config():
    import bench
    cfg, n, m = bench.WORKLOADS["c2"]
    assert set(bench.config_of("c2", n, m, 1)) == set(bench.config_of("c2", n, m, 8))
    assert bench.config_of("c2", n, m, 1)["workload"] == bench.WORKLOAD_TEXT["c2"]
