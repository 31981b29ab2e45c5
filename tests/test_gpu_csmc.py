"""GPU parity: conditional bootstrap filters (conditionalSampler on the LGSS example), plain LGSS filter, compareNCestimates and
ancestral lines, against the reference fixtures and the oracle port on identical draws."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
PAR = (0.9, 1.0, 0.0, 1.0)
MODES = (H.MULTINOMIAL, H.RESIDUAL, H.STRATIFIED, H.SYSTEMATIC)


@pytest.fixture(scope="module")
def smc():
    import rcppsmc_b200 as s
    assert s.device_count() >= 1
    return s


def _compare_history(got, k, exp):
    assert abs(got["logNC"][k] - exp["logNC"]) <= 1e-9 * abs(exp["logNC"])
    assert np.array_equal(got["resampled"][k], exp["resampled"])
    assert np.array_equal(got["ancestors"][k], exp["ancestors"])       # bit-exact index work
    assert np.array_equal(got["refidx"][k], exp["refidx"])
    assert np.array_equal(got["values"][k], exp["values"])             # IEEE mul/add on identical draws and ancestors
    np.testing.assert_allclose(got["logw"][k], exp["logw"], rtol=0, atol=1e-11)


def test_reference_fixture_chain(smc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "csmc_kat.npz"))
    y, par, N, K = g["y"], tuple(g["par"]), int(g["N"]), int(g["ncases"])
    modes = [int(g[f"mode{k}"]) for k in range(K)]
    ref = np.stack([g[f"ref{k}"] for k in range(K)])
    z = np.stack([g[f"z{k}"] for k in range(K)])
    U = np.stack([g[f"U{k}"] for k in range(K)])
    got = smc.conditionalBPF(y, N, par, modes, ref, noise=z, uniforms=U, history=True)   # ONE chain: the fixture is one sampler object
    for k in range(K):
        _compare_history(got, k, {key: g[f"{key}{k}"] for key in ("logNC", "resampled", "ancestors", "refidx", "values", "logw")})
        lines, lv = smc.ancestor_lines(got["ancestors"][k], got["values"][k])
        assert np.array_equal(lines, g[f"lines{k}"]) and np.array_equal(lv, g[f"line_values{k}"])


@pytest.mark.parametrize("N,T", [(100, 25), (1000, 40), (777, 12), (20000, 10)])
def test_chains_match_oracle_on_identical_draws(smc, N, T):
    x, y = H.sim_lgss(T, N)
    rng = np.random.default_rng(N + T)
    chains, rpc = 3, 4
    runs = chains * rpc
    modes = [MODES[(r + r // rpc) % 4] for r in range(runs)]            # every chain in a different order
    ref = x + rng.standard_normal((runs, T))
    z, U = rng.standard_normal((runs, T, N)), rng.random((runs, T, N + 2))
    got = smc.conditionalBPF(y, N, PAR, modes, ref, runs_per_chain=rpc, noise=z, uniforms=U, history=True)
    n_res = 0
    for c in range(chains):
        refidx = np.zeros(T, dtype=np.uint32)
        for r in range(c * rpc, (c + 1) * rpc):
            exp = H.o_csmc_run(y, N, modes[r], PAR, ref[r], refidx, z[r], U[r])
            _compare_history(got, r, exp)
            n_res += int(exp["resampled"].sum())
    assert n_res >= runs


@pytest.mark.parametrize("mode", MODES)
def test_plain_lgss_filter_matches_oracle(smc, mode):
    N, T = 3000, 40
    _, y = H.sim_lgss(T, 9)
    rng = np.random.default_rng(mode)
    per = {H.SYSTEMATIC: 1, H.STRATIFIED: N + 1}.get(mode, N)
    z, U = rng.standard_normal((T, N)), rng.random((T, per))
    got = smc.lgssBPF(y, N, PAR, mode, noise=z, uniforms=U)
    exp, flags = H.o_bpf_lgss(y, N, mode, PAR, z, U, want_flags=True)
    assert flags.sum() >= 3
    assert abs(got - exp) <= 1e-6 * abs(exp)


def test_ancestor_lines_large(smc):
    T, N = 40, 1 << 18
    rng = np.random.default_rng(2)
    anc = np.sort(rng.integers(0, N, size=(T, N)), axis=1).astype(np.uint32)   # resampling-like maps
    anc[::3] = np.arange(N, dtype=np.uint32)                                  # steps without resampling
    val = rng.standard_normal((T, N))
    lines, lv = smc.ancestor_lines(anc, val)
    el, ev = H.o_ancestor_lines(anc, val)
    assert np.array_equal(lines, el) and np.array_equal(lv, ev)
    assert np.array_equal(smc.ancestor_lines(anc), el)
    with pytest.raises(smc.SmcError):
        bad = anc.copy(); bad[0, 0] = N
        smc.ancestor_lines(bad)


def _kalman_loglik(y, par):
    phi, q, x0, r = par
    m, P, ll = x0, 0.0, 0.0
    for t in range(y.size):
        m, P = phi * m, phi * phi * P + q
        S = P + r
        ll += -0.5 * (np.log(2 * np.pi * S) + (y[t] - m) ** 2 / S)
        K = P / S
        m, P = m + K * (y[t] - m), (1 - K) * P
    return ll


def test_compareNCestimates_native(smc):
    T, N, S = 30, 500, 48
    x, y = H.sim_lgss(T, 4)
    rng = np.random.default_rng(1)
    ref = (x[:, None] + 0.3 * rng.standard_normal((T, S)))
    a = smc.compareNCestimates(y, N, S, dict(phi=PAR[0], varStateEvol=PAR[1], stateInit=PAR[2], varObs=PAR[3]), ref, seed=7)
    b = smc.compareNCestimates(y, N, S, PAR, ref, seed=7, independent_sims=True)
    assert a["smcOut"].shape == (S, 4) and a["csmcOut"].shape == (S, 4)
    assert np.array_equal(a["smcOut"], b["smcOut"])
    truth = _kalman_loglik(y, PAR)
    # plain bootstrap filters: exp(logNC) is unbiased for the likelihood; logNC is within Monte-Carlo error of the Kalman value
    for c in range(4):
        v = a["smcOut"][:, c]
        assert abs(v.mean() - truth) <= 4 * v.std(ddof=1) / np.sqrt(S) + 0.05
    # conditional filters agree in distribution with the oracle port run the same way (its own RNG): 3 MC standard errors
    refidx = np.zeros(T, dtype=np.uint32)
    exp = np.zeros((S, 4))
    for i in range(S):
        for c, mode in enumerate(MODES):
            exp[i, c] = H.o_csmc_run(y, N, mode, PAR, ref[:, i], refidx, seed=1000 + 4 * i + c)["logNC"]
    for got in (a["csmcOut"], b["csmcOut"]):
        for c in range(4):
            se = np.sqrt(got[:, c].var(ddof=1) / S + exp[:, c].var(ddof=1) / S)
            assert abs(got[:, c].mean() - exp[:, c].mean()) <= 3 * se + 1e-9, (c, got[:, c].mean(), exp[:, c].mean(), se)
