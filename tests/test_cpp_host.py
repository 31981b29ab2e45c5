"""The C++ host side above the C ABI (include/smcb200.hpp: the reference's exported function names on std::vector): builds a
small program against the header and the shared library.  CPU: it must compile, link and fail loudly without a device.
GPU: its results equal the Python mirror's for the same seeds (same library underneath)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_smoke")


def _build():
    import rcppsmc_b200 as smc
    smc.lib()   # makes sure the library exists
    libdir = os.path.join(ROOT, "rcppsmc_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_smoke.cpp"),
           "-L" + libdir, "-l:libsmcb200.so", "-Wl,-rpath," + libdir, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_host_builds_and_fails_loudly_without_a_device():
    import rcppsmc_b200 as smc
    exe = _build()
    if smc.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 12, (r.returncode, r.stderr)          # 10 + SMCB200_ERR_CUDA
    assert "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_host_matches_the_python_mirror():
    sys.path.insert(0, ROOT)
    import rcppsmc_b200 as smc
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = {ln.split()[0]: ln.split() for ln in r.stdout.strip().splitlines()}
    y = 3.0 * np.sin(0.7 * np.arange(24.0)) + 0.1 * np.arange(24.0)
    a = smc.pfNonlinBS(y, 4096, resample_mode=smc.SYSTEMATIC, seed=7)
    assert np.array_equal(np.array([float(v) for v in lines["pfNonlinBS"][2:]]), a["mean"])
    b = smc.blockpfGaussianOpt(y, 512, 3, seed=5)
    assert float(lines["blockpfGaussianOpt"][2]) == b["logNC"] and abs(float(lines["blockpfGaussianOpt"][4]) - 1.0) < 1e-9
    p = smc.nonLinPMMH(y, 512, 3, seed=1)
    assert float(lines["nonLinPMMH"][2]) == p["loglike"][0] and int(lines["nonLinPMMH"][4]) == 4
    c = smc.compareNCestimates(y, 200, 2, (0.9, 1.0, 0.0, 1.0), np.zeros((24, 2)), seed=1)
    assert float(lines["compareNCestimates"][2]) == c["smcOut"][0, 0] and float(lines["compareNCestimates"][4]) == c["csmcOut"][0, 0]
    assert lines["argument"][-1] == "1"


def test_rcpp_glue_is_well_formed_against_the_stand_in_headers():
    # glue/smcb200_rcpp_glue.cpp holds the eight [[Rcpp::export]] bodies a maintainer drops into the R package.  R / Rcpp /
    # Armadillo are absent here, so this is a syntax / type check against oracle/shim's stand-ins, not a real Rcpp build.
    src = os.path.join(ROOT, "glue", "smcb200_rcpp_glue.cpp")
    text = open(src).read()
    for name in ("pfNonlinBS_impl", "pfLineartBS_impl", "LinReg_impl", "LinRegLA_impl", "LinRegLA_adapt_impl", "blockpfGaussianOpt_impl",
                 "nonLinPMMH_impl", "compareNCestimates_imp"):
        assert text.count("// [[Rcpp::export]]\n") == 8 and (" " + name + "(") in text, name
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
