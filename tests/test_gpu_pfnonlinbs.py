"""GPU parity, model level: pfNonlinBS through the C ABI vs the oracle on identical draws, and vs Philox runs."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # BASELINE.json north_star: logNC / ESS traces within 1e-6 relative in fp64 on identical draws


@pytest.fixture(scope="module")
def smc():
    import rcppsmc_b200 as s
    assert s.device_count() >= 1
    return s


def _inject_case(N, T, mode, seed):
    rng = np.random.default_rng(seed)
    _, y = H.sim_nonlin(T, seed)
    z = rng.standard_normal(N * T)
    U = rng.random({H.SYSTEMATIC: T, H.STRATIFIED: T * (N + 1)}.get(mode, T * N))
    return y, z, U


@pytest.mark.parametrize("N,T", [(1000, 50), (1, 5), (2, 7), (37, 20), (4099, 30), (1 << 15, 12)])
@pytest.mark.parametrize("mode", [H.SYSTEMATIC, H.STRATIFIED, H.MULTINOMIAL, H.RESIDUAL])
def test_traces_match_oracle_on_identical_draws(smc, N, T, mode):
    y, z, U = _inject_case(N, T, mode, seed=N + T + mode)
    got = smc.pfNonlinBS(y, N, resample_mode=mode, threshold=1.01 * N, noise=z, uniforms=U, full=True)
    exp = H.o_pfnonlinbs(y, N, mode, 1.01 * N, z=z, U=U)
    assert np.array_equal(got["resampled"], exp["resampled"])
    np.testing.assert_allclose(got["logNC"], exp["logNC"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=RTOL)
    np.testing.assert_allclose(got["mean"], exp["mean"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["sd"], exp["sd"], rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("thr", [0.5, 0.9])
def test_ess_triggered_resampling_matches_oracle(smc, thr):
    # threshold < 1 is a fraction of N (sampler.h:885-893); some steps resample, some do not
    N, T = 2000, 40
    y, z, U = _inject_case(N, T, H.SYSTEMATIC, seed=5)
    got = smc.pfNonlinBS(y, N, resample_mode=H.SYSTEMATIC, threshold=thr, noise=z, uniforms=U, full=True)
    exp = H.o_pfnonlinbs(y, N, H.SYSTEMATIC, thr, z=z, U=U, by_step=True)  # the ABI indexes injected uniforms by step
    assert 0 < exp["resampled"].sum() < T
    assert np.array_equal(got["resampled"], exp["resampled"])
    np.testing.assert_allclose(got["logNC"], exp["logNC"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=RTOL)
    np.testing.assert_allclose(got["mean"], exp["mean"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["sd"], exp["sd"], rtol=RTOL, atol=1e-9)


def test_ancestors_of_last_step_bit_exact(smc):
    N, T = 3000, 9
    y, z, U = _inject_case(N, T, H.SYSTEMATIC, seed=11)
    f = smc.PfNonlinBS(N, T, seed=1)
    f.set_data(y)
    f.set_noise(z, U)
    f.run(H.SYSTEMATIC, 1.01 * N)
    idx = f.read_ancestors()
    exp = H.o_pfnonlinbs(y, N, H.SYSTEMATIC, 1.01 * N, z=z, U=U, want_idx=True)
    assert np.array_equal(idx, exp["idx"][T - 1])
    f.close()


def test_philox_runs_agree_with_oracle_within_monte_carlo_error(smc):
    # native RNG: filtering means must agree with the CPU reference path within 3 Monte Carlo standard errors
    N, T, R = 4000, 25, 24
    _, y = H.sim_nonlin(T, 3)
    gm = np.array([smc.pfNonlinBS(y, N, resample_mode=H.SYSTEMATIC, seed=100 + r)["mean"] for r in range(R)])
    om = np.array([H.o_pfnonlinbs(y, N, H.SYSTEMATIC, 1.01 * N, seed=500 + r)["mean"] for r in range(R)])
    se = np.sqrt(gm.var(axis=0, ddof=1) / R + om.var(axis=0, ddof=1) / R)
    zscore = np.abs(gm.mean(axis=0) - om.mean(axis=0)) / se
    assert (zscore < 3.0).mean() >= 0.95 and zscore.max() < 4.5


def test_reference_default_mode_multinomial_statistics(smc):
    # the drop-in default of pfNonlinBS_impl (MULTINOMIAL, threshold 1.01 N; src/pfnonlinbs.cpp:62) under Philox:
    # filtering means agree with SYSTEMATIC runs of the oracle within Monte Carlo error
    N, T, R = 4000, 20, 20
    _, y = H.sim_nonlin(T, 4)
    gm = np.array([smc.pfNonlinBS(y, N, seed=300 + r)["mean"] for r in range(R)])   # defaults = reference behaviour
    om = np.array([H.o_pfnonlinbs(y, N, H.MULTINOMIAL, 1.01 * N, seed=700 + r)["mean"] for r in range(R)])
    se = np.sqrt(gm.var(axis=0, ddof=1) / R + om.var(axis=0, ddof=1) / R)
    zscore = np.abs(gm.mean(axis=0) - om.mean(axis=0)) / se
    assert (zscore < 3.0).mean() >= 0.9 and zscore.max() < 4.5


def test_philox_run_is_deterministic_and_seed_sensitive(smc):
    _, y = H.sim_nonlin(30, 9)
    a = smc.pfNonlinBS(y, 10000, resample_mode=H.SYSTEMATIC, seed=42, full=True)
    b = smc.pfNonlinBS(y, 10000, resample_mode=H.SYSTEMATIC, seed=42, full=True)
    c = smc.pfNonlinBS(y, 10000, resample_mode=H.SYSTEMATIC, seed=43, full=True)
    for k in ("mean", "sd", "logNC", "ESS"):
        assert np.array_equal(a[k], b[k])
    assert not np.array_equal(a["mean"], c["mean"])


def test_full_size_properties(smc):
    # BASELINE headline size: 2^24 particles (shortened T): size-independent invariants
    N, T = 1 << 24, 6
    _, y = H.sim_nonlin(T, 1)
    f = smc.PfNonlinBS(N, T, seed=3)
    f.set_data(y)
    f.run(H.SYSTEMATIC, 1.01 * N)
    out = f.read()
    assert np.all(out["resampled"] == 1)
    assert np.all((out["ESS"] > 1.0) & (out["ESS"] <= N))
    assert np.all(np.isfinite(out["logNC"])) and np.all(np.diff(out["logNC"]) < 0.0)
    idx = f.read_ancestors().astype(np.int64)
    cnt = np.bincount(idx, minlength=N)
    assert cnt.sum() == N
    # reference in-place map: every parent with children keeps its own slot, others are filled in order
    assert np.all(idx[cnt > 0] == np.flatnonzero(cnt > 0))
    zero_slots = np.flatnonzero(cnt == 0)
    assert np.all(np.diff(idx[zero_slots]) >= 0)
    # small-N run with the same data agrees statistically
    small = smc.pfNonlinBS(y, 1 << 16, resample_mode=H.SYSTEMATIC, seed=8)
    assert np.all(np.abs(small["mean"] - out["mean"]) < 0.5)
    f.close()


@pytest.mark.parametrize("mode", [H.RESIDUAL, H.MULTINOMIAL])
def test_repeated_runs_on_one_handle_are_bit_identical(smc, mode):
    # same seed, same handle, 2^20 particles: every run must reproduce the first one bit for bit
    T, N = 40, 1 << 20
    _, y = H.sim_nonlin(60, 1)
    f = smc.PfNonlinBS(N, T, seed=1)
    f.set_data(y[:T])
    first = None
    for _ in range(4):
        f.run(mode)
        f.sync()
        out = f.read()
        idx = f.read_ancestors()
        if first is None:
            first = (out["logNC"].copy(), idx.copy())
            assert np.bincount(idx.astype(np.int64), minlength=N).sum() == N
        else:
            assert np.array_equal(out["logNC"], first[0]) and np.array_equal(idx, first[1])
    f.close()


@pytest.mark.parametrize("thr", [1.01, 0.5])
def test_mispredicted_weight_shift_takes_the_cold_path_and_still_matches(smc, monkeypatch, thr):
    # k_step stores exp(lw - c) against a shift c predicted by the previous step; an unusable c (here: off by e^1000 through
    # the experiment switch, so that every shifted weight underflows) must be detected and repaired from the stored log-weights
    N, T = 5000, 12
    y, z, U = _inject_case(N, T, H.SYSTEMATIC, seed=21)
    monkeypatch.setenv("SMCB200_DBG_STEP", "32")
    f = smc.PfNonlinBS(N, T, seed=1)     # the switch is read when the sampler is built
    monkeypatch.delenv("SMCB200_DBG_STEP")
    f.set_data(y)
    f.set_noise(z, U)
    f.run(H.SYSTEMATIC, thr * N if thr > 1 else thr)
    got = f.read()
    f.close()
    exp = H.o_pfnonlinbs(y, N, H.SYSTEMATIC, thr * N if thr > 1 else thr, z=z, U=U, by_step=True)
    assert np.array_equal(got["resampled"], exp["resampled"])
    np.testing.assert_allclose(got["logNC"], exp["logNC"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=RTOL)
    np.testing.assert_allclose(got["mean"], exp["mean"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["sd"], exp["sd"], rtol=RTOL, atol=1e-9)
