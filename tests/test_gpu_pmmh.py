"""GPU parity: nonLinPMMH -- inner filter log-likelihood and the outer MH chain vs the oracle on identical draws."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smc():
    import rcppsmc_b200 as s
    assert s.device_count() >= 1
    return s


def _data(T, seed):
    # R/nonLinPMMH.R:8: simNonlin(len, var_init = 5, var_evol = 10, var_obs = 1, cosSeqOffset = 0)
    return H.sim_nonlin(T, seed, var_init=5.0, cos_offset=0.0)[1]


@pytest.mark.parametrize("mode,thr", [(H.SYSTEMATIC, 1.01), (H.MULTINOMIAL, 0.5), (H.STRATIFIED, 0.5), (H.RESIDUAL, 0.5)])
@pytest.mark.parametrize("N,T", [(2000, 60), (1 << 16, 25)])
def test_inner_filter_loglik_matches_oracle(smc, mode, thr, N, T):
    y = _data(T, N)
    rng = np.random.default_rng(mode + N)
    per = {H.SYSTEMATIC: 1, H.STRATIFIED: N + 1}.get(mode, N)
    z, U = rng.standard_normal(N * T), rng.random(T * per)
    threshold = thr if thr < 1 else thr * N
    got = smc.nonLinPMMH(y, N, 0, resample_mode=mode, threshold=threshold, filter_noise=z, filter_uniforms=U)
    exp = H.o_pmmh_filter(y, N, 10.0, 10.0, mode, threshold, z=z, U=U)
    assert abs(got["loglike"][0] - exp) <= 1e-6 * abs(exp)
    assert got["samples_sigv"][0] == 10.0 and got["samples_sigw"][0] == 10.0


def test_outer_chain_matches_oracle_on_identical_draws(smc):
    N, T, M = 400, 30, 25
    y = _data(T, 1)
    rng = np.random.default_rng(9)
    innov = (0.15 * rng.standard_normal(M), 0.08 * rng.standard_normal(M), rng.random(M))
    z, U = rng.standard_normal((M + 1) * N * T), rng.random((M + 1) * N * T)
    got = smc.nonLinPMMH(y, N, M, innovations=innov, filter_noise=z, filter_uniforms=U)
    exp = H.o_pmmh(y, N, M, H.MULTINOMIAL, 0.5, innov, z, U)
    assert np.array_equal(got["MH_acceptance_rate"], exp["MH_acceptance_rate"])
    assert np.array_equal(got["samples_sigv"], exp["samples_sigv"]) and np.array_equal(got["samples_sigw"], exp["samples_sigw"])
    np.testing.assert_allclose(got["loglike"], exp["loglike"], rtol=1e-6)
    np.testing.assert_allclose(got["logprior"], exp["logprior"], rtol=1e-12)
    assert 0 < got["MH_acceptance_rate"][-1] < 1


def test_negative_proposals_are_rejected(smc):
    # a proposal with a negative standard deviation gives NaN and must be rejected
    y = _data(20, 2)
    innov = (np.array([-30.0, 0.1]), np.array([0.0, -40.0]), np.array([0.5, 0.5]))
    out = smc.nonLinPMMH(y, 500, 2, innovations=innov)
    assert np.all(out["samples_sigv"] == 10.0) and np.all(out["samples_sigw"] == 10.0)
    assert np.all(out["MH_acceptance_rate"] == 0.0)


def test_native_chain_moves_towards_the_posterior_and_reports_progress(smc):
    y = _data(100, 3)
    seen = []
    out = smc.nonLinPMMH(y, 1000, 300, seed=5, msg_freq=100, progress=lambda *a: seen.append(a))
    assert [s[0] for s in seen] == [100, 200, 300]
    assert 0.02 < out["MH_acceptance_rate"][-1] < 0.9
    assert out["samples_sigw"][-50:].mean() < 9.6           # started at 10 with steps of s.d. 0.08; the data were simulated with sigw = 1
    assert out["loglike"][-1] > out["loglike"][0]
    again = smc.nonLinPMMH(y, 1000, 20, seed=5)
    assert np.array_equal(again["samples_sigv"], out["samples_sigv"][:21])   # deterministic given the seed


def test_graph_replayed_filters_reproduce_the_eager_chain(smc):
    # chains of >= 64 iterations replay one captured CUDA graph per inner filter (seed and proposal travel through the
    # control block); a short chain runs the same filters eagerly: the common prefix must be bit-identical
    y = _data(40, 5)
    short = smc.nonLinPMMH(y, 3000, 20, seed=11)
    long = smc.nonLinPMMH(y, 3000, 80, seed=11)
    for k in ("samples_sigv", "samples_sigw", "loglike", "logprior", "MH_acceptance_rate"):
        assert np.array_equal(long[k][:21], short[k]), k
    assert len(np.unique(long["loglike"])) > 5
