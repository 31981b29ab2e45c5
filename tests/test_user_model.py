"""The plug-in boundary: a user model written outside the engine against include/smcb200/sampler.cuh only --
smc::sampler<Model>, smc::adaptMethods, smc::staticModelAdapt, HistoryType::AL, IntegratePathSampling, IterateBack under the
reference's names -- compiled out of tree with nvcc.  tests/cpp/user_model.cu checks the device run against a CPU loop of
the same __host__ __device__ functors and against a quadrature of the evidence; here we build it and run it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "user_model")
SRC = os.path.join(ROOT, "tests", "cpp", "user_model.cu")
HDRS = [os.path.join(ROOT, "include", "smcb200", "sampler.cuh")] + [
    os.path.join(ROOT, "rcppsmc_b200", "csrc", f) for f in ("generic.cuh", "resample.cuh", "common.cuh", "philox.cuh", "fastmath.cuh")]


def _build():
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(f) for f in [SRC] + HDRS):
        return EXE
    cmd = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_user_model_compiles_out_of_tree_and_fails_loudly_without_a_device():
    import rcppsmc_b200 as smc
    exe = _build()
    if smc.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 12 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("particles", [4096, 65536])
def test_user_model_matches_cpu_loop_and_quadrature(particles):
    exe = _build()
    r = subprocess.run([exe, str(particles)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "user_model OK" in r.stdout
