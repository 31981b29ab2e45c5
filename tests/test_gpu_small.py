"""GPU: the single-launch filter for small populations (csrc/small.cuh) against the oracle on injected draws, and against the
kernel-per-step path (SMCB200_SMALL_MAX=0) on the device's own Philox draws: same traces, same ancestors."""
import os

import numpy as np
import pytest

import helpers as H
import rcppsmc_b200 as smc

pytestmark = pytest.mark.gpu

MODES = [smc.MULTINOMIAL, smc.RESIDUAL, smc.STRATIFIED, smc.SYSTEMATIC]


def _run(N, T, y, mode, thr, seed, small):
    old = os.environ.get("SMCB200_SMALL_MAX")
    os.environ["SMCB200_SMALL_MAX"] = str(1 << 18) if small else "0"
    try:
        f = smc.PfNonlinBS(N, T, seed=seed)
        f.set_data(y)
        f.run(mode, thr)
        f.sync()
        out = f.read()
        anc = f.read_ancestors()
        launches = f.launches()
        f.close()
    finally:
        if old is None:
            os.environ.pop("SMCB200_SMALL_MAX", None)
        else:
            os.environ["SMCB200_SMALL_MAX"] = old
    return out, anc, launches


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("N,thr", [(1000, None), (1000, 0.5), (4098, 0.5), (65536, None), (65536, 0.5), (200001, 0.5)])
def test_small_path_equals_kernel_per_step_path(mode, N, thr):
    T = 30
    _, y = H.sim_nonlin(T, 5)
    thr_abs = 1.01 * N if thr is None else thr
    a, anc_a, la = _run(N, T, y, mode, thr_abs, 11, True)
    b, anc_b, lb = _run(N, T, y, mode, thr_abs, 11, False)
    assert la == 1 and lb > 1          # one launch for the whole filter vs several per step
    assert np.array_equal(a["resampled"], b["resampled"])
    # same draws, same definitions: only the order of the partial sums differs
    for k in ("logNC", "ESS", "mean", "sd"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-9, atol=1e-9, err_msg=k)
    if a["resampled"][-1]:
        # the last step's ancestors: identical up to a handful of boundary flips caused by the last-bit difference of the normaliser
        assert np.mean(anc_a != anc_b) < 1e-3


@pytest.mark.parametrize("mode", [smc.STRATIFIED, smc.SYSTEMATIC])
@pytest.mark.parametrize("N", [1000, 4099, 65536])
def test_small_path_vs_oracle_on_injected_draws(mode, N):
    T = 20
    rng = np.random.default_rng(N + mode)
    _, y = H.sim_nonlin(T, 2)
    z = rng.standard_normal(N * T)
    U = rng.random(T) if mode == smc.SYSTEMATIC else rng.random(T * (N + 1))
    for thr in (1.01 * N, 0.5 * N):
        got = smc.pfNonlinBS(y, N, resample_mode=mode, threshold=thr, noise=z, uniforms=U, full=True)
        exp = H.o_pfnonlinbs(y, N, mode, thr, z=z, U=U, by_step=True)   # the ABI indexes injected uniforms by step
        assert np.array_equal(got["resampled"], exp["resampled"])
        np.testing.assert_allclose(got["logNC"], exp["logNC"], rtol=1e-6)
        np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=1e-6)
        np.testing.assert_allclose(got["mean"], exp["mean"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got["sd"], exp["sd"], rtol=1e-6, atol=1e-9)


def test_small_path_mispredicted_shift():
    """an outlier observation moves the weight level by more than e^250 between two steps: the cold path redoes the sums"""
    T, N = 12, 5000
    _, y = H.sim_nonlin(T, 3)
    y[6] = 400.0
    a, _, _ = _run(N, T, y, smc.SYSTEMATIC, 0.5 * N, 3, True)
    b, _, _ = _run(N, T, y, smc.SYSTEMATIC, 0.5 * N, 3, False)
    assert np.all(np.isfinite(a["logNC"]))
    np.testing.assert_allclose(a["logNC"], b["logNC"], rtol=1e-9)
    np.testing.assert_allclose(a["ESS"], b["ESS"], rtol=1e-6)
