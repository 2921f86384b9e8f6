"""The C++11 drop-in header (include/perturb/satellite_batch.hpp) compiled like user code and,
on a GPU, run against the golden fixtures of the reference."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import load_ver_golden
from perturb_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "iss_example.cpp")
REF_INCLUDE = "/root/reference/include"


SHARDED_SRC = os.path.join(ROOT, "tests", "cpp", "sharded_example.cpp")


def _compile(tmp_path, extra=(), src=SRC, name="iss_example"):
    exe = str(tmp_path / name)
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           *extra, src, "-o", exe, "-L" + libdir, "-lperturb_gpu", "-lperturb_benchtools",
           "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_header_compiles_as_cxx11_user_code(tmp_path):
    exe = _compile(tmp_path)
    if not os.path.exists("/dev/nvidia0"):
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 2 and "no CPU fallback" in p.stdout  # fails loudly, no fallback


@pytest.mark.skipif(not os.path.isdir(REF_INCLUDE), reason="reference tree not present")
def test_header_compiles_against_the_reference_headers(tmp_path):
    """Integration mode: perturb's own JulianDate / StateVector / TwoLineElement types."""
    subprocess.check_call(
        ["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Wextra", "-DPERTURB_BATCH_USE_PERTURB_HPP",
         "-I" + REF_INCLUDE, "-I" + os.path.join(ROOT, "include"), SRC])


@pytest.mark.gpu
def test_cpp_example_matches_reference(tmp_path):
    G = load_ver_golden()
    exe = _compile(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    init = [ln.split() for ln in out if ln.startswith("init")]
    assert [int(x[3]) for x in init] == [0, 0, 0, 7]  # last TLE is too short: INVALID_TLE
    assert float(init[0][5]) == G["iss_epoch"][0] and float(init[0][6]) == G["iss_epoch"][1]
    jd = [ln.split() for ln in out if ln.startswith("jd ")]
    r = np.array([float(x) for x in jd[0][5:8]])
    v = np.array([float(x) for x in jd[0][9:12]])
    assert int(jd[0][3]) == 0
    assert np.abs(r - G["iss_r"]).max() < 1e-6 and np.abs(v - G["iss_v"]).max() < 1e-9
    # rejected TLE: MEAN_MOTION on propagate and the caller's StateVector left untouched
    assert int(jd[3][3]) == 2 and [float(x) for x in jd[3][5:8]] == [-1.0, -1.0, -1.0]
    # sv.epoch is always set: to the requested JulianDate / to epoch() + mins / 1440, bit for bit
    ep = [ln.split() for ln in out if ln.startswith("jdepoch")]
    want_frac = (26.535 + 59 * 60.0 + 1 * 3600.0) / 86400.0
    assert len(ep) == 4 and all(float(x[2]) == 2459652.5 and float(x[3]) == want_frac for x in ep)
    mep = [ln.split() for ln in out if ln.startswith("minsepoch")]
    assert len(mep) == 12 and all(x[3] == "1" for x in mep)
    mins = [ln.split() for ln in out if ln.startswith("mins 1 ")]
    # satellite 00005 with opsmode 'i' (public API); near-Earth, so identical to the 'a' rows
    rows = G["full_i"]
    for ln in mins:
        t = float(ln[2])
        ref = rows[(rows[:, 0] == 0) & (rows[:, 1] == t)][0]
        got_r = np.array([float(x) for x in ln[6:9]])
        got_v = np.array([float(x) for x in ln[10:13]])
        assert int(ln[4]) == int(ref[2])
        assert np.abs(got_r - ref[3:6]).max() < 1e-6 and np.abs(got_v - ref[6:9]).max() < 1e-9


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_sharded_header_compiles_as_cxx11_user_code(tmp_path):
    exe = _compile(tmp_path, src=SHARDED_SRC, name="sharded_example")
    if not os.path.exists("/dev/nvidia0"):
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 2 and "no CPU fallback" in p.stdout  # fails loudly, no fallback


def test_cpp_range_plan_matches_the_python_plan(tmp_path):
    """ShardedSatelliteBatch::plan_ranges / estimated_cost (C++) against perturb_b200.sharding:
    same class thresholds, cost-balanced contiguous cuts (host logic, no GPU)."""
    import perturb_b200 as pb
    n = 4000
    tl, _ = pb.parse_tle_blocks(*pb.synth_catalog(3, n), n)
    src = tmp_path / "plan.cpp"
    src.write_text(r"""
#include <cstdio>
#include <vector>
#include <perturb/satellite_batch.hpp>
int main() {
    std::vector<double> cost;
    double mm, ecc;
    while (std::scanf("%lf %lf", &mm, &ecc) == 2) {
        perturb::TwoLineElement t = perturb::TwoLineElement();
        t.mean_motion = mm; t.eccentricity = ecc;
        cost.push_back(perturb::ShardedSatelliteBatch::estimated_cost(t));
    }
    for (double c : cost) std::printf("%.0f\n", c);
    for (std::size_t k : {1u, 3u, 8u}) {
        std::printf("cuts");
        for (std::size_t c : perturb::ShardedSatelliteBatch::plan_ranges(cost, k)) std::printf(" %zu", c);
        std::printf("\n");
    }
    return 0;
}
""")
    exe = str(tmp_path / "plan")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                           "-L" + libdir, "-lperturb_gpu", "-Wl,-rpath," + libdir])
    feed = "".join("%.17g %.17g\n" % (t.mean_motion, t.eccentricity) for t in tl)
    out = subprocess.run([exe], input=feed, capture_output=True, text=True, check=True).stdout.split("\n")
    cost = np.array([float(x) for x in out[:n]])
    want = np.array(pb.sharding.CLASS_COST)[pb.estimate_class(tl)]
    assert np.array_equal(cost, want)
    for line, k in zip(out[n:n + 3], (1, 3, 8)):
        cuts = [int(x) for x in line.split()[1:]]
        assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == k + 1 and cuts == sorted(cuts)
        sums = np.array([want[a:b].sum() for a, b in zip(cuts, cuts[1:])])
        assert sums.max() - sums.min() <= 2 * 2097.0


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [1, 3, 8])
def test_cpp_sharded_batch_equals_the_unsharded_one(tmp_path, shards):
    exe = _compile(tmp_path, src=SHARDED_SRC, name="sharded_example")
    out = subprocess.run([exe, str(shards)], capture_output=True, text=True, check=True).stdout
    assert "init_mismatches 0" in out, out
    assert "planes_equal 1 codes_equal 1" in out, out
    assert "records_equal 1 record_codes_equal 1 status 0" in out, out
    assert "device_gather_equal 1 device_gather_codes_equal 1 status 0" in out, out
    # ShardedSatelliteBatch::from_tles + propagate(JulianDate*, m, StateVector*, Sgp4Error*): the
    # unsharded drop-in's bits, failed evaluations left untouched, same range plan
    drop = [ln for ln in out.splitlines() if ln.startswith("dropin_equal")][0].split()
    assert drop[1] == "1" and drop[3] == "1" and drop[5] == "0", out
    assert drop[7] == drop[9] and drop[11] == "1", out
    cuts = [int(x) for x in out.splitlines()[0].split("cuts")[1].split()]
    assert len(cuts) == shards and cuts[-1] == 3000


JULIAN_SRC = os.path.join(ROOT, "tests", "cpp", "julian_dates.cpp")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_julian_date_and_datetime_equal_the_reference(tmp_path):
    """The stand-alone perturb::JulianDate / DateTime (include/perturb/batch_types.hpp) against
    the reference's (perturb.cpp:40-117 over jday_SGP4 / invjday_SGP4): conversions both ways
    for 24 776 time points of 1901-2099, normalisation, offsets and ordering print the same
    text, digit for digit, as the same program built on the reference's headers and sources
    (golden hash made by tests/golden/make_cpp_goldens.sh; compared live when the reference
    tree is present)."""
    import hashlib
    exe = str(tmp_path / "julian_ours")
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror",
                           "-I" + os.path.join(ROOT, "include"), JULIAN_SRC, "-o", exe])
    ours = subprocess.run([exe], capture_output=True, check=True).stdout
    want_hash, want_lines = open(os.path.join(ROOT, "tests", "golden", "julian_dates.sha256")).read().split()
    assert len(ours.splitlines()) == int(want_lines)
    assert hashlib.sha256(ours).hexdigest() == want_hash
    if os.path.isdir(REF_INCLUDE):
        ref_exe = str(tmp_path / "julian_ref")
        src = "/root/reference/src/"
        subprocess.check_call(["g++", "-std=c++11", "-w", "-DPERTURB_BATCH_USE_PERTURB_HPP",
                               "-I" + REF_INCLUDE, "-I" + os.path.join(ROOT, "include"), JULIAN_SRC,
                               src + "perturb.cpp", src + "sgp4.cpp", src + "tle.cpp", "-o", ref_exe])
        ref = subprocess.run([ref_exe], capture_output=True, check=True).stdout
        assert ours == ref
    # the reference's own round-trip bar (tests/test_perturb.cpp:109-129): jd exact, frac to 1e-12
    for ln in ours.decode().splitlines()[:4776:97]:
        f = ln.split()
        jd, fr = float(f[1]), float(f[2])
        y, mo, d, h, mi, s = int(f[4]), int(f[5]), int(f[6]), int(f[7]), int(f[8]), float(f[9])
        jd2 = 367.0 * y - np.floor(7 * (y + np.floor((mo + 9) / 12.0)) * 0.25) + np.floor(275 * mo / 9.0) + d + 1721013.5
        assert jd2 == jd and abs((s + mi * 60.0 + h * 3600.0) / 86400.0 - fr) < 1e-12


TLE_SRC = os.path.join(ROOT, "tests", "cpp", "tle_members.cpp")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_twolineelement_parse_equals_the_reference_member_for_member(tmp_path):
    """The stand-alone perturb::TwoLineElement::parse (include/perturb/batch_types.hpp over
    perturb_gpu_parse_tle) against the reference's sscanf reader (src/tle.cpp:41-173): 2 036
    well-formed TLEs and 30 000 corrupted ones give the same TLEParseError and, when accepted,
    the same 21 members digit for digit -- including the lines the reference rejects because a
    blank field shifts its tokens (golden hash made by tests/golden/make_cpp_goldens.sh from the
    reference build; compared live when the reference tree is present)."""
    import hashlib
    from helpers import tle_parser_feed
    feed = tle_parser_feed()
    exe = _compile(tmp_path, src=TLE_SRC, name="tle_members")
    ours = subprocess.run([exe], input=feed, capture_output=True, check=True).stdout
    gold = open(os.path.join(ROOT, "tests", "golden", "tle_members.sha256")).read().split("\n")
    assert len(ours.splitlines()) == int(gold[1])
    assert hashlib.sha256(ours).hexdigest() == gold[0]
    # every error class occurs in the feed (counts recorded next to the hash)
    for code in range(5):
        assert ours.count(b"err %d\n" % code) > 3000
    if os.path.isdir(REF_INCLUDE):
        ref_exe = str(tmp_path / "tle_members_ref")
        src = "/root/reference/src/"
        subprocess.check_call(["g++", "-std=c++11", "-w", "-DPERTURB_BATCH_USE_PERTURB_HPP",
                               "-I" + REF_INCLUDE, "-I" + os.path.join(ROOT, "include"), TLE_SRC,
                               src + "perturb.cpp", src + "sgp4.cpp", src + "tle.cpp", "-o", ref_exe])
        ref = subprocess.run([ref_exe], input=feed, capture_output=True, check=True).stdout
        assert ours == ref


# the reference's developer preset (CMakePresets.json:57-62 "flags-unix"), as errors
REFERENCE_WARNING_FLAGS = (
    "-Wall -Wextra -Wpedantic -Wconversion -Wsign-conversion -Wcast-qual -Wcast-align -Wshadow "
    "-Wformat=2 -Wformat-overflow -Wformat-truncation -Wundef -Wdouble-promotion "
    "-Wstrict-overflow=5 -Wno-maybe-uninitialized -Wno-unused-result").split()


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
@pytest.mark.parametrize("src", ["iss_example.cpp", "sharded_example.cpp", "julian_dates.cpp",
                                 "tle_members.cpp"])
def test_headers_are_clean_under_the_reference_developer_flags(src):
    """The drop-in headers compile as C++11 without a single warning under the warning set the
    reference builds itself with in developer mode."""
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-fsyntax-only", "-Werror",
                           *REFERENCE_WARNING_FLAGS, "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", src)])
