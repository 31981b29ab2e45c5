"""CPU: the oracle (oracle/smc_oracle.c) against the committed fixtures generated from the compiled reference."""
import os

import numpy as np
import pytest

import helpers as H


def _load(name):
    return np.load(os.path.join(H.GOLDEN, name))


def test_lse_ess_bit_exact_vs_reference_fixtures():
    g = _load("lse_kat.npz")
    for i in range(int(g["ncases"])):
        lw = g[f"lw{i}"]
        assert H.o_lse(lw) == float(g[f"lse{i}"])
        assert H.o_ess(lw) == float(g[f"ess{i}"])
    assert float(g["lse0"]) == -0.07985233853818718  # SURVEY.md App. B / skeleton example


def test_resample_indices_bit_exact_vs_reference_fixtures():
    g = _load("resample_kat.npz")
    n_counts = n_w = n_lw = 0
    for k in range(int(g["ncases"])):
        mode, lw, U, counts, w, idx = int(g[f"mode{k}"]), g[f"lw{k}"], g[f"U{k}"], g[f"counts{k}"], g[f"w{k}"], g[f"idx{k}"]
        if counts.size:
            got, _ = H.o_resample_w(mode, None, counts=counts)
            n_counts += 1
        elif w.size:
            got, _ = H.o_resample_w(mode, w, U)
            n_w += 1
        else:
            got, _ = H.o_resample_lw(mode, lw, U)
            n_lw += 1
        assert np.array_equal(got, idx), f"case {k} (mode {mode}, n={idx.size})"
    assert n_counts >= 10 and n_w >= 20 and n_lw >= 20


def test_appendix_b_vectors():
    g = _load("resample_kat.npz")
    assert np.array_equal(g["idx0"], [0, 1, 1, 3, 3, 5, 6, 5])   # SYSTEMATIC, U = 0.37
    assert np.array_equal(g["idx1"], [0, 1, 1, 3, 5, 5, 6, 6])   # STRATIFIED
    assert np.array_equal(g["idx2"], [1, 1, 1, 3, 3, 5, 6, 5])   # MULTINOMIAL counts 0 3 0 2 0 2 1 0


def test_pfnonlinbs_traces_bit_exact_vs_reference_fixtures():
    g = _load("pfnonlinbs_kat.npz")
    for k in range(int(g["ncases"])):
        N, T, mode, thr = int(g[f"N{k}"]), int(g[f"T{k}"]), int(g[f"mode{k}"]), float(g[f"thr{k}"])
        r = H.o_pfnonlinbs(g[f"y{k}"], N, mode, thr, z=g[f"z{k}"], U=g[f"U{k}"])
        assert np.array_equal(r["resampled"], g[f"resampled{k}"])
        for key in ("mean", "sd", "logNC"):
            assert np.array_equal(r[key], g[f"{key}{k}"]), f"case {k} {key}"
        assert np.array_equal(r["ESS"][1:], g[f"ESS{k}"][1:])  # the reference does not expose the t=0 ESS


def test_quantised_weights_make_cumsum_order_independent():
    rng = np.random.default_rng(3)
    w = H.random_weights(rng, 100000, 4.0)
    fwd = np.cumsum(w)
    # any association order gives the same bits on the 2^-52 grid
    blocks = np.add.accumulate(w.reshape(100, 1000), axis=1)
    offs = np.concatenate([[0.0], np.cumsum(blocks[:, -1])[:-1]])
    assert np.array_equal((blocks + offs[:, None]).ravel(), fwd)
    assert abs(fwd[-1] - 1.0) < 1e-11


def test_native_multinomial_and_residual_definitions():
    rng = np.random.default_rng(4)
    n = 5000
    w = H.random_weights(rng, n, 2.0)
    U = rng.random(n)
    idx, cnt = H.o_resample_w(H.MULTINOMIAL, w, U)
    assert cnt.sum() == n and np.array_equal(np.bincount(idx, minlength=n), cnt)
    # inverse CDF: counts follow the weights
    assert abs((cnt * np.arange(n)).sum() / n - (w * np.arange(n)).sum()) < 5 * n / np.sqrt(n)
    idx, cnt = H.o_resample_w(H.RESIDUAL, w, U)
    assert cnt.sum() == n
    assert np.all(cnt >= np.floor(n * w - 1e-9).astype(int))
    assert np.all(cnt <= np.floor(n * w).astype(int) + (n - np.floor(n * w).sum()))


def test_pflineart_traces_bit_exact_vs_reference_fixtures():
    g = _load("pflineart_kat.npz")
    for k in range(int(g["ncases"])):
        N, mode, thr = int(g[f"N{k}"]), int(g[f"mode{k}"]), float(g[f"thr{k}"])
        r = H.o_pflineart(g[f"data{k}"], N, mode, thr, z=g[f"z{k}"], U=g[f"U{k}"])
        assert np.array_equal(r["resampled"], g[f"resampled{k}"])
        assert 0 < r["resampled"].sum() <= r["resampled"].size
        for key in ("Xm", "Xv", "Ym", "Yv", "logNC"):
            assert np.array_equal(r[key], g[f"{key}{k}"]), f"case {k} {key}"
        assert np.array_equal(r["ESS"][1:], g[f"ESS{k}"][1:])


def test_blockpf_bit_exact_vs_reference_fixtures():
    g = _load("blockpf_kat.npz")
    resampled = 0
    for k in range(int(g["ncases"])):
        r = H.o_blockpf(g[f"y{k}"], int(g[f"N{k}"]), int(g[f"lag{k}"]), g[f"z{k}"], g[f"U{k}"])
        assert r["logNC"] == float(g[f"logNC{k}"]), f"case {k}"
        assert np.array_equal(r["weight"], g[f"weight{k}"]) and np.array_equal(r["values"], g[f"values{k}"]), f"case {k}"
        resampled += int(r["resampled"].sum())
    assert resampled >= 3   # the fixtures exercise the path copy on resampling


def test_conditional_sampler_chain_bit_exact_vs_reference_fixtures():
    g = _load("csmc_kat.npz")
    y, par, N = g["y"], tuple(g["par"]), int(g["N"])
    refidx = np.zeros(y.size, dtype=np.uint32)
    for k in range(int(g["ncases"])):
        o = H.o_csmc_run(y, N, int(g[f"mode{k}"]), par, g[f"ref{k}"], refidx, g[f"z{k}"], g[f"U{k}"])
        assert o["logNC"] == float(g[f"logNC{k}"]), k
        for key in ("ancestors", "values", "logw", "resampled", "refidx"):
            assert np.array_equal(o[key], g[f"{key}{k}"]), (k, key)
        lines, lv = H.o_ancestor_lines(o["ancestors"], o["values"])
        assert np.array_equal(lines, g[f"lines{k}"]) and np.array_equal(lv, g[f"line_values{k}"])
