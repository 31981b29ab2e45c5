"""GPU parity, model level: pfLineartBS (4-d state, Student-t weights, ESS-triggered resampling) vs the oracle."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def smc():
    import rcppsmc_b200 as s
    assert s.device_count() >= 1
    return s


def _uniforms_by_step(rng, mode, N, T):
    per = {H.SYSTEMATIC: 1, H.STRATIFIED: N + 1}.get(mode, N)
    return rng.random(T * per), per


@pytest.mark.parametrize("N,T", [(1000, 100), (3, 8), (4099, 40), (1 << 15, 20)])
@pytest.mark.parametrize("mode", [H.RESIDUAL, H.SYSTEMATIC, H.STRATIFIED, H.MULTINOMIAL])
@pytest.mark.parametrize("thr", [0.5, 1.01])
def test_traces_match_oracle_on_identical_draws(smc, N, T, mode, thr):
    rng = np.random.default_rng(N + T + mode)
    data = H.sim_lineart(T, N + mode)
    z = rng.standard_normal(4 * N * T)
    U, per = _uniforms_by_step(rng, mode, N, T)
    threshold = thr if thr < 1 else thr * N
    got = smc.pfLineartBS(data, N, resample_mode=mode, threshold=threshold, noise=z, uniforms=U, full=True)
    exp = H.o_pflineart(data, N, mode, threshold, z=z, U=U, by_step=True)  # the ABI indexes injected uniforms by step
    assert np.array_equal(got["resampled"], exp["resampled"])
    np.testing.assert_allclose(got["logNC"], exp["logNC"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(got["ESS"], exp["ESS"], rtol=RTOL)
    for k in ("Xm", "Ym"):
        np.testing.assert_allclose(got[k], exp[k], rtol=RTOL, atol=1e-8)
    for k in ("Xv", "Yv"):
        np.testing.assert_allclose(got[k], exp[k], rtol=1e-5, atol=1e-10)


def test_reference_default_and_callback(smc):
    # defaults = reference behaviour (RESIDUAL, 0.5); the per-step callback mirrors `if (usef) fun(Xm, Ym)`
    T, N = 30, 5000
    data = H.sim_lineart(T, 1)
    calls = []
    out = smc.pfLineartBS(data, N, usef=True, fun=lambda xm, ym: calls.append((xm.copy(), ym.copy())), seed=5)
    assert len(calls) == T - 1
    last_xm, last_ym = calls[-1]
    np.testing.assert_allclose(last_xm, out["Xm"], rtol=1e-12)
    np.testing.assert_allclose(last_ym, out["Ym"], rtol=1e-12)
    assert np.all(calls[0][0][2:] == 0.0)          # entries beyond the current step are still zero
    same = smc.pfLineartBS(data, N, seed=5)
    assert np.array_equal(same["Xm"], out["Xm"]) and np.array_equal(same["Yv"], out["Yv"])


def test_philox_filter_tracks_the_oracle_statistically(smc):
    T, N, R = 40, 4000, 16
    data = H.sim_lineart(T, 2)
    g = np.array([smc.pfLineartBS(data, N, seed=10 + r)["Xm"] for r in range(R)])
    o = np.array([H.o_pflineart(data, N, H.RESIDUAL, 0.5, seed=90 + r)["Xm"] for r in range(R)])
    se = np.sqrt(g.var(axis=0, ddof=1) / R + o.var(axis=0, ddof=1) / R)
    zs = np.abs(g.mean(axis=0) - o.mean(axis=0)) / se
    assert (zs < 3.0).mean() >= 0.9 and zs.max() < 4.5
