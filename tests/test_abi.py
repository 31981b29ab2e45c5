"""CPU: the C-ABI shared library loads, exports every symbol include/smcb200.h declares, and refuses to compute
without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb200_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import rcppsmc_b200 as smc
    L = smc.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/smcb200.h but not exported"
    assert L.smcb200_version() >= 100


def test_no_torch_types_in_header():
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    assert "torch" not in text and "at::" not in text and "std::" not in text


def test_philox_known_answers_host_side():
    # Random123 known-answer vectors for philox4x32-10
    import rcppsmc_b200 as smc
    assert [hex(v) for v in smc.philox4x32([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in smc.philox4x32([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in smc.philox4x32([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_compute_fails_loudly_without_device():
    import rcppsmc_b200 as smc
    if smc.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.log_sum_exp(np.zeros(4))
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.resample(smc.SYSTEMATIC, np.ones(4) / 4, [0.5])
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.pfNonlinBS(np.zeros(5), 100, resample_mode=smc.SYSTEMATIC)
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.blockpfGaussianOpt(np.zeros(5), 100, 2)
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.compareNCestimates(np.zeros(5), 100, 2, (0.9, 1.0, 0.0, 1.0), np.zeros((5, 2)))
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.conditionalBPF(np.zeros(5), 100, (0.9, 1.0, 0.0, 1.0), [smc.SYSTEMATIC], np.zeros((1, 5)))
    with pytest.raises(smc.SmcError, match="no CUDA device"):
        smc.ancestor_lines(np.zeros((3, 4), dtype=np.uint32))


def test_argument_errors_are_reported_through_the_status_channel():
    # invalid arguments are rejected before any device work (status SMCB200_ERR_INVALID + smcb200_last_error text),
    # the C-ABI counterpart of smc::exception / Rcpp::stop in the reference drivers
    import rcppsmc_b200 as smc
    with pytest.raises(smc.SmcError, match="lag >= 1"):
        smc.blockpfGaussianOpt(np.zeros(5), 100, 0)
    with pytest.raises(smc.SmcError, match="particles >= 1"):
        smc.blockpfGaussianOpt(np.zeros(5), 0, 2)
    with pytest.raises(smc.SmcError, match="unknown resampling mode"):
        smc.conditionalBPF(np.zeros(5), 100, (0.9, 1.0, 0.0, 1.0), [7], np.zeros((1, 5)))
    with pytest.raises(smc.SmcError, match="T >= 2"):
        smc.conditionalBPF(np.zeros(1), 100, (0.9, 1.0, 0.0, 1.0), [smc.SYSTEMATIC], np.zeros((1, 1)))
    with pytest.raises(smc.SmcError, match="out of range"):
        smc.ancestor_lines(np.full((2, 3), 9, dtype=np.uint32))
    with pytest.raises(ValueError):
        smc.compareNCestimates(np.zeros(5), 100, 2, (0.9, 1.0, 0.0, 1.0), np.zeros((4, 2)))


def test_product_does_not_touch_the_oracle():
    # the product tree must not reference oracle/ (or any CPU fallback) anywhere
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rcppsmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "smc_oracle" not in text and "libsmcref" not in text, f
