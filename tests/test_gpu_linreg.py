"""GPU parity: likelihood-annealing SMC samplers (LinRegLA_adapt, LinRegLA) vs the oracle on identical draws, and the
reference's known log-evidence (man/LinReg.Rd:78-80: -309.9 for model 1, -301.4 for model 2)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smc():
    import rcppsmc_b200 as s
    assert s.device_count() >= 1
    return s


def _draws(seed, n_norm, n_unif):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n_norm), rng.random(n_unif)


@pytest.mark.parametrize("N,resampTol,tempTol", [(1000, 0.5, 0.9), (513, 0.7, 0.8), (4096, 0.5, 0.95)])
def test_adaptive_sampler_matches_oracle_on_identical_draws(smc, N, resampTol, tempTol):
    data = H.radiata()[:, [0, 1]]
    z, U = _draws(N, 40_000_000 if N > 2000 else 12_000_000, 14_000_000 if N > 2000 else 4_000_000)
    exp = H.o_linreg_la(data, N, True, None, resampTol, tempTol, z=z, U=U)
    got = smc.LinRegLA_adapt(data, N, resampTol, tempTol, noise=z, uniforms=U)
    L = exp["Temps"].size
    assert got["Temps"].size == L
    np.testing.assert_allclose(got["Temps"], exp["Temps"], rtol=1e-8)
    assert np.array_equal(got["mcmcRepeats"], exp["mcmcRepeats"])
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=1e-6)
    for k in ("logNC_standard", "logNC_ps_rect", "logNC_ps_trap", "logNC_ps_trap2"):
        assert abs(got[k] - exp[k]) <= 1e-6 * abs(exp[k]), k
    np.testing.assert_allclose(got["acceptProb"], exp["acceptProb"], rtol=1e-12)
    np.testing.assert_allclose(got["theta"], exp["theta"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got["loglike"], exp["loglike"], rtol=1e-9)
    np.testing.assert_allclose(got["logprior"], exp["logprior"], rtol=1e-9)
    np.testing.assert_allclose(got["Weights"], exp["Weights"], rtol=1e-7, atol=1e-300)


def test_fixed_schedule_sampler_matches_oracle_on_identical_draws(smc):
    data = H.radiata()[:, [0, 1]]
    N, L = 800, 30
    temps = (np.arange(L) / (L - 1.0)) ** 5
    z, U = _draws(7, 2 * N + 3 * N * 10 * L + 10, 3 * N + N * 10 * L + L + 10)
    exp = H.o_linreg_la(data, N, False, temps, z=z, U=U, max_temps=L)
    got = smc.LinRegLA(data, temps, N, noise=z, uniforms=U)
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=1e-6)
    for k in ("logNC_standard", "logNC_ps_rect", "logNC_ps_trap", "logNC_ps_trap2"):
        assert abs(got[k] - exp[k]) <= 1e-6 * abs(exp[k]), k
    np.testing.assert_allclose(got["theta"], exp["theta"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got["loglike"], exp["loglike"], rtol=1e-9)
    np.testing.assert_allclose(got["Weights"], exp["Weights"], rtol=1e-7, atol=1e-300)


@pytest.mark.parametrize("col,known", [(1, -309.9), (2, -301.4)])
def test_known_log_evidence_under_native_rng(smc, col, known):
    # BASELINE config 3 size: 2^18 particles, adaptive temperatures + MCMC rejuvenation, Philox draws
    data = H.radiata()[:, [0, col]]
    out = smc.LinRegLA_adapt(data, 1 << 18, 0.5, 0.9, seed=11)
    assert out["Temps"][0] == 0.0 and out["Temps"][-1] == 1.0 and np.all(np.diff(out["Temps"]) > 0)
    for k in ("logNC_standard", "logNC_ps_trap2"):
        assert abs(out[k] - known) < 0.15, (k, out[k])
    assert abs(out["logNC_ps_trap"] - known) < 0.35     # plain trapezoid: discretisation bias of the adaptive grid
    assert abs(out["logNC_ps_rect"] - known) < 1.5      # the rectangle rule is biased by O(step)
    assert np.all(out["ESS"] > 0.45 * (1 << 18))
    # posterior mean of (alpha, beta) under the final weights: OLS-like values for the radiata regression
    w = out["Weights"][:, -1]
    alpha = float((w * out["theta"][:, 0, -1]).sum() / w.sum())
    assert 2950 < alpha < 3050


def test_data_annealing_sampler_matches_oracle_on_identical_draws(smc):
    # LinReg_impl: one observation per step, native multinomial resampling, 10 MH repeats per step
    data = H.radiata()[:, [0, 1]]
    N = 600
    rng = np.random.default_rng(21)
    z, U = rng.standard_normal(2 * N + 3 * N * 10 * 42), rng.random(3 * N + N * 10 * 42)
    Ud = rng.random(42 * N)
    exp = H.o_linreg(data, N, z=z, U=U, Udraw=Ud)
    got = smc.LinReg(data, N, noise=z, uniforms=U, resample_uniforms=Ud)
    assert abs(got["logNC"] - exp["logNC"]) <= 1e-6 * abs(exp["logNC"])
    np.testing.assert_allclose(got["theta"], exp["theta"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got["weights"], exp["weights"], rtol=1e-7)


def test_data_annealing_known_log_evidence(smc):
    data = H.radiata()[:, [0, 1]]
    out = smc.LinReg(data, 1 << 17, seed=4)
    assert abs(out["logNC"] - (-309.9)) < 0.25       # man/LinReg.Rd:78-80
    assert abs(out["weights"].sum() - 1.0) < 1e-9


def test_adaptive_schedule_longer_than_the_first_history_capacity(smc):
    """the wrappers size the history for 64 temperatures first and repeat with the caller's capacity when the schedule is longer"""
    Data = H.radiata()[:, :2]
    r = smc.LinRegLA_adapt(Data, 600, 0.5, 0.997, seed=3)
    L = r["Temps"].size
    assert L > 64 and r["Temps"][-1] == 1.0 and np.all(np.diff(r["Temps"]) > 0)
    assert r["theta"].shape == (600, 3, L) and np.isfinite(r["logNC_standard"])
