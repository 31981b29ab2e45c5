"""CPU: the oracle restatement against the UNMODIFIED reference compiled from /root/reference (oracle/_ref).
Skipped where the compiled reference is absent."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.skipif(not H.ref_available(), reason="oracle/_ref/libsmcref.so not built (no /root/reference here)")


@pytest.mark.parametrize("n", [1, 2, 7, 100, 1001, 20000])
def test_lse_ess(n):
    rng = np.random.default_rng(n)
    lw = 6.0 * rng.standard_normal(n) + 50.0
    assert H.o_lse(lw) == H.r_lse(lw)
    assert H.o_ess(lw) == H.r_ess(lw)


@pytest.mark.parametrize("n", [1, 2, 3, 64, 999, 30000])
@pytest.mark.parametrize("mode", [H.SYSTEMATIC, H.STRATIFIED])
def test_resample_from_logweights(n, mode):
    rng = np.random.default_rng(7 * n + mode)
    for spread in (0.3, 4.0, 15.0):
        lw = spread * rng.standard_normal(n)
        U = rng.random(1 if mode == H.SYSTEMATIC else n + 1)
        ridx, rval = H.r_resample(mode, lw, uniforms=U)
        oidx, _ = H.o_resample_lw(mode, lw, U)
        assert np.array_equal(ridx, oidx)
        assert np.array_equal(rval, ridx.astype(float))  # replication really copied value[idx]


@pytest.mark.parametrize("n", [8, 513, 25000])
@pytest.mark.parametrize("mode", [H.SYSTEMATIC, H.STRATIFIED])
def test_resample_identical_gridded_weights(n, mode):
    # the bit-exactness surface: the same normalised weights (2^-52 grid) and uniforms into both
    rng = np.random.default_rng(n + mode)
    w = H.random_weights(rng, n, 5.0)
    U = rng.random(1 if mode == H.SYSTEMATIC else n + 1)
    ridx, _ = H.r_resample(mode, np.zeros(n), uniforms=U, weights_override=w)
    oidx, _ = H.o_resample_w(mode, w, U)
    assert np.array_equal(ridx, oidx)


@pytest.mark.parametrize("n", [8, 1000, 40000])
def test_count_to_index_map(n):
    rng = np.random.default_rng(n)
    for total in (n, n - 3 if n > 3 else n):
        counts = rng.multinomial(total, rng.dirichlet(np.full(n, 0.2))).astype(np.int32)
        ridx, _ = H.r_resample(H.MULTINOMIAL, np.zeros(n), multinomial=counts)
        oidx, _ = H.o_resample_w(H.MULTINOMIAL, None, counts=counts)
        assert np.array_equal(ridx, oidx)


@pytest.mark.parametrize("N,T,mode,thr", [(1000, 50, H.SYSTEMATIC, 1010.0), (1000, 50, H.STRATIFIED, 1010.0),
                                           (500, 40, H.SYSTEMATIC, 0.5), (300, 30, H.STRATIFIED, 0.7), (1, 5, H.SYSTEMATIC, 1.01)])
def test_pfnonlinbs_traces(N, T, mode, thr):
    rng = np.random.default_rng(N + T)
    _, y = H.sim_nonlin(T, N)
    z, U = rng.standard_normal(N * T), rng.random(T * (N + 1))
    r = H.r_pfnonlinbs(y, N, mode, thr, z, U)
    o = H.o_pfnonlinbs(y, N, mode, thr, z=z, U=U)
    assert np.array_equal(r["resampled"], o["resampled"])
    for key in ("mean", "sd", "logNC"):
        assert np.array_equal(r[key], o[key])
    assert np.array_equal(r["ESS"][1:], o["ESS"][1:])


@pytest.mark.parametrize("N,T,mode,thr", [(500, 40, H.SYSTEMATIC, 0.5), (300, 30, H.STRATIFIED, 0.5), (200, 25, H.SYSTEMATIC, 201.0),
                                           (2, 6, H.STRATIFIED, 0.9)])
def test_pflineart_traces(N, T, mode, thr):
    # ESS-triggered resampling (threshold 0.5 N as in the reference driver) and always-resample
    rng = np.random.default_rng(3 * N + T)
    data = H.sim_lineart(T, N)
    z, U = rng.standard_normal(4 * N * T), rng.random(T * (N + 1))
    r = H.r_pflineart(data, N, mode, thr, z, U)
    o = H.o_pflineart(data, N, mode, thr, z=z, U=U)
    assert np.array_equal(r["resampled"], o["resampled"])
    for key in ("Xm", "Xv", "Ym", "Yv", "logNC"):
        assert np.array_equal(r[key], o[key]), key
    assert np.array_equal(r["ESS"][1:], o["ESS"][1:])


def test_linreg_la_adapt_and_fixed_schedule_bit_exact():
    # tempering samplers: CESS bisection, empirical covariance + Cholesky, adaptive MCMC repeats, RAM history, path sampling
    import ctypes
    data = H.radiata()[:, [0, 1]]
    flat = np.ascontiguousarray(data.T.reshape(-1))
    N, maxT = 250, 200
    rng = np.random.default_rng(8)
    z, U = rng.standard_normal(3_000_000), rng.random(3_000_000)
    R = H.ref()
    R.smcref_linreg_la_adapt.restype = H.L
    th = np.zeros(maxT * 3 * N); ll = np.zeros(maxT * N); lp = np.zeros(maxT * N); w = np.zeros(maxT * N)
    ess, temps, reps, nc = np.zeros(maxT), np.zeros(maxT), np.zeros(maxT), np.zeros(4)
    H.r_clear(); H.r_push(normals=z, uniforms=U)
    Lr = R.smcref_linreg_la_adapt(H.dp(flat), H.L(42), H.L(N), H.D(0.5), H.D(0.9), H.L(maxT), H.dp(th), H.dp(ll), H.dp(lp), H.dp(w),
                                  H.dp(ess), H.dp(temps), H.dp(reps), H.dp(nc))
    o = H.o_linreg_la(data, N, True, None, 0.5, 0.9, z=z, U=U, max_temps=maxT)
    assert Lr == o["Temps"].size and Lr > 5
    assert np.array_equal(temps[:Lr], o["Temps"]) and np.array_equal(reps[:Lr], o["mcmcRepeats"]) and np.array_equal(ess[:Lr], o["ESS"])
    assert np.array_equal(th[:Lr * 3 * N].reshape(Lr, 3, N).transpose(2, 1, 0), o["theta"])
    assert np.array_equal(w[:Lr * N].reshape(Lr, N).T, o["Weights"])
    assert list(nc) == [o["logNC_standard"], o["logNC_ps_rect"], o["logNC_ps_trap"], o["logNC_ps_trap2"]]
    # fixed schedule (LinRegLA_impl)
    Lt = 20
    sched = (np.arange(Lt) / (Lt - 1.0)) ** 5
    th = np.zeros(Lt * 3 * N); ll = np.zeros(Lt * N); lp = np.zeros(Lt * N); w = np.zeros(Lt * N); ess = np.zeros(Lt)
    H.r_clear(); H.r_push(normals=z, uniforms=U)
    assert R.smcref_linreg_la(H.dp(flat), H.L(42), H.dp(sched), H.L(Lt), H.L(N), H.dp(th), H.dp(ll), H.dp(lp), H.dp(w), H.dp(ess), H.dp(nc)) == 0
    o = H.o_linreg_la(data, N, False, sched, z=z, U=U, max_temps=Lt)
    assert np.array_equal(ess, o["ESS"]) and np.array_equal(w.reshape(Lt, N).T, o["Weights"])
    assert list(nc) == [o["logNC_standard"], o["logNC_ps_rect"], o["logNC_ps_trap"], o["logNC_ps_trap2"]]


def test_linreg_data_annealing_log_evidence_agrees_statistically():
    # LinReg_impl resamples through rmultinom (no uniform-level parity): compare the evidence estimates
    data = H.radiata()[:, [0, 1]]
    flat = np.ascontiguousarray(data.T.reshape(-1))
    N, reps = 400, 6
    R = H.ref()
    ref_nc, ora_nc = [], []
    for r in range(reps):
        import ctypes as _c
        H.r_clear()
        nc = np.zeros(1)
        R.smcref_seed(_c.c_uint64(100 + r))
        assert R.smcref_linreg(H.dp(flat), H.L(42), H.L(N), None, None, H.dp(nc)) == 0
        ref_nc.append(nc[0])
        ora_nc.append(H.o_linreg(data, N, seed=200 + r)["logNC"])
    ref_nc, ora_nc = np.array(ref_nc), np.array(ora_nc)
    assert abs(ref_nc.mean() - (-309.9)) < 0.6 and abs(ora_nc.mean() - (-309.9)) < 0.6
    se = np.sqrt(ref_nc.var(ddof=1) / reps + ora_nc.var(ddof=1) / reps)
    assert abs(ref_nc.mean() - ora_nc.mean()) < 4 * se + 0.05


def test_pmmh_inner_filter_bit_exact():
    import ctypes
    R, O = H.ref(), H.oracle()
    R.smcref_pmmh_filter.restype = ctypes.c_double
    _, y = H.sim_nonlin(40, 5, var_init=5.0, cos_offset=0.0)
    N, T = 500, 40
    rng = np.random.default_rng(0)
    for mode in (H.SYSTEMATIC, H.STRATIFIED):
        per = 1 if mode == H.SYSTEMATIC else N + 1
        z, U = rng.standard_normal(N * T), rng.random(T * per)
        H.r_clear(); H.r_push(normals=z, uniforms=U)
        a = R.smcref_pmmh_filter(H.dp(H.f64(y)), H.L(T), H.L(N), H.D(9.3), H.D(1.2), mode, H.D(1.01 * N))
        assert a == H.o_pmmh_filter(y, N, 9.3, 1.2, mode, 1.01 * N, z=z, U=U)


@pytest.mark.parametrize("N,T,lag,sd", [(50, 40, 4, 3.0), (128, 20, 2, 5.0), (9, 15, 30, 2.0), (31, 9, 1, 4.0)])
def test_blockpf_bit_exact(N, T, lag, sd):
    y = H.sim_lg(T, N + T, sd)
    rng = np.random.default_rng(N * T + lag)
    z, U = rng.standard_normal((T, lag, N)), rng.random(T)
    o = H.o_blockpf(y, N, lag, z, U)
    r = H.r_blockpf(y, N, lag, z, U[o["resampled"] == 1])
    assert o["logNC"] == r["logNC"]
    assert np.array_equal(o["weight"], r["weight"]) and np.array_equal(o["values"], r["values"])


LGSS_PAR = (0.9, 1.0, 0.0, 1.0)   # phi, varStateEvol, stateInit, varObs


def test_conditional_sampler_chain_bit_exact():
    # a chain of runs on ONE conditionalSampler object, as the reference driver does (cSMCexamples.cpp:81-125): the
    # reference-index state carries over between runs and modes
    T, N = 16, 48
    x, y = H.sim_lgss(T, 1)
    rng = np.random.default_rng(5)
    H.r_csmc_open(y, N, LGSS_PAR)
    refidx = np.zeros(T, dtype=np.uint32)
    seen = set()
    for sim in range(3):
        reft = x + rng.standard_normal(T)
        for mode in (H.MULTINOMIAL, H.RESIDUAL, H.STRATIFIED, H.SYSTEMATIC):
            z, U = rng.standard_normal((T, N)), rng.random((T, N + 2))
            o = H.o_csmc_run(y, N, mode, LGSS_PAR, reft, refidx, z, U)
            r = H.r_csmc_run(mode, reft, T, N, z, H.csmc_uniform_queue(mode, U, o, N))
            assert o["logNC"] == r["logNC"]
            for key in ("ancestors", "values", "logw", "refidx", "resampled"):
                assert np.array_equal(o[key], r[key]), (sim, mode, key)
            assert o["resampled"].sum() >= 2
            lines, lv = H.o_ancestor_lines(o["ancestors"], o["values"])     # GetALineInd / GetALineSpace for every particle
            assert np.array_equal(lines, r["lines"]) and np.array_equal(lv, r["line_values"])
            seen.add(tuple(r["refidx"]))
    assert len(seen) > 3   # the carried state really varies


@pytest.mark.parametrize("mode", [H.STRATIFIED, H.SYSTEMATIC])
def test_plain_lgss_filter_bit_exact(mode):
    T, N = 30, 100
    _, y = H.sim_lgss(T, 3)
    rng = np.random.default_rng(mode)
    per = 1 if mode == H.SYSTEMATIC else N + 1
    z, U = rng.standard_normal((T, N)), rng.random((T, per))
    H.r_csmc_open(y, N, LGSS_PAR)
    v, flags = H.o_bpf_lgss(y, N, mode, LGSS_PAR, z, U, want_flags=True)
    assert flags.sum() >= 3
    assert v == H.r_bpf_lgss(mode, z, U[flags == 1].reshape(-1))


def test_randomised_configurations_blockpf_and_conditional_chain():
    # breadth: many small random shapes (odd N, short series, lag >= T, every scheme order) stay bit-identical
    rng = np.random.default_rng(2026)
    for _ in range(40):
        N, T, lag = int(rng.integers(2, 90)), int(rng.integers(1, 25)), int(rng.integers(1, 12))
        y = H.sim_lg(T, int(rng.integers(1 << 30)), float(rng.uniform(1.0, 30.0)))
        z, U = rng.standard_normal((T, lag, N)), rng.random(T)
        o = H.o_blockpf(y, N, lag, z, U)
        r = H.r_blockpf(y, N, lag, z, U[o["resampled"] == 1])
        assert o["logNC"] == r["logNC"] and np.array_equal(o["values"], r["values"]) and np.array_equal(o["weight"], r["weight"]), (N, T, lag)
    for _ in range(15):
        N, T = int(rng.integers(8, 70)), int(rng.integers(3, 20))
        par = (float(rng.uniform(0.5, 0.99)), float(rng.uniform(0.3, 2.0)), float(rng.normal()), float(rng.uniform(0.3, 2.0)))
        x, y = H.sim_lgss(T, int(rng.integers(1 << 30)), par[0], par[1], par[3], par[2])
        H.r_csmc_open(y, N, par)
        refidx = np.zeros(T, dtype=np.uint32)
        for mode in rng.permutation([H.MULTINOMIAL, H.RESIDUAL, H.STRATIFIED, H.SYSTEMATIC, H.RESIDUAL, H.SYSTEMATIC]):
            mode = int(mode)
            reft = x + rng.standard_normal(T)
            z, U = rng.standard_normal((T, N)), rng.random((T, N + 2))
            o = H.o_csmc_run(y, N, mode, par, reft, refidx, z, U)
            r = H.r_csmc_run(mode, reft, T, N, z, H.csmc_uniform_queue(mode, U, o, N))
            assert o["logNC"] == r["logNC"], (N, T, mode)
            for key in ("ancestors", "values", "logw", "refidx", "resampled"):
                assert np.array_equal(o[key], r[key]), (N, T, mode, key)


def test_randomised_configurations_filters():
    # breadth for the two bootstrap filters: random sizes, thresholds on both sides of 1, both exact-parity schemes
    rng = np.random.default_rng(77)
    for k in range(24):
        N, T = int(rng.integers(1, 400)), int(rng.integers(1, 30))
        mode = H.SYSTEMATIC if k % 2 else H.STRATIFIED
        thr = float(rng.choice([0.3, 0.5, 0.9, 1.01 * N, 0.75 * N]))
        z, U = rng.standard_normal(4 * N * T), rng.random(T * (N + 1))
        if k % 3:
            _, y = H.sim_nonlin(T, int(rng.integers(1 << 30)))
            r, o = H.r_pfnonlinbs(y, N, mode, thr, z[:N * T], U), H.o_pfnonlinbs(y, N, mode, thr, z=z[:N * T], U=U)
            keys = ("mean", "sd", "logNC")
        else:
            data = H.sim_lineart(T, int(rng.integers(1 << 30)))
            r, o = H.r_pflineart(data, N, mode, thr, z, U), H.o_pflineart(data, N, mode, thr, z=z, U=U)
            keys = ("Xm", "Xv", "Ym", "Yv", "logNC")
        assert np.array_equal(r["resampled"], o["resampled"]), (k, N, T, mode, thr)
        for key in keys:
            assert np.array_equal(r[key], o[key]), (k, N, T, mode, thr, key)
