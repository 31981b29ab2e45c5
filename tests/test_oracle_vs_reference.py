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
