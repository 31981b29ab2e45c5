"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled against oracle/shim
(oracle/_ref/libsmcref.so, built by `make -C oracle ref` from /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors of its own, so these fixtures are outputs of
the reference code itself on fixed, injected randomness.  They pin the oracle on machines where /root/reference
is absent (the GPU box)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402


def blockpf_fixture():
    """blockpfGaussianOpt_impl outputs of the compiled reference on injected draws (src/blockpfgaussianopt.cpp:41-71)."""
    rng = np.random.default_rng(20251020)
    out = {}
    k = 0
    for (N, T, lag, sd) in ((64, 30, 3, 3.0), (100, 25, 5, 4.0), (33, 12, 1, 3.0), (17, 10, 20, 3.0), (8, 2, 3, 1.0), (8, 1, 3, 1.0)):
        y = H.sim_lg(T, 300 + k, sd)
        z, U = rng.standard_normal((T, lag, N)), rng.random(T)
        o = H.o_blockpf(y, N, lag, z, U)            # only to learn which steps consume a uniform; the values stored are the reference's
        r = H.r_blockpf(y, N, lag, z, U[o["resampled"] == 1])
        out[f"N{k}"] = np.array(N); out[f"lag{k}"] = np.array(lag); out[f"y{k}"] = y; out[f"z{k}"] = z; out[f"U{k}"] = U
        for key in ("weight", "values", "logNC"):
            out[f"{key}{k}"] = np.asarray(r[key])
        k += 1
    out["ncases"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "blockpf_kat.npz"), **out)


def csmc_fixture():
    """A chain of conditional bootstrap-filter runs on ONE reference conditionalSampler object (the driver's usage pattern,
    src/cSMCexamples.cpp:81-125) on injected draws: logNC, AL history, reference-index state, ancestral lines."""
    rng = np.random.default_rng(20251021)
    par = (0.9, 1.0, 0.0, 1.0)
    T, N = 20, 64
    x, y = H.sim_lgss(T, 11)
    H.r_csmc_open(y, N, par)
    refidx = np.zeros(T, dtype=np.uint32)
    out = {"y": y, "par": np.array(par), "N": np.array(N)}
    k = 0
    for sim in range(2):
        reft = x + rng.standard_normal(T)
        for mode in (H.MULTINOMIAL, H.RESIDUAL, H.STRATIFIED, H.SYSTEMATIC):
            z, U = rng.standard_normal((T, N)), rng.random((T, N + 2))
            o = H.o_csmc_run(y, N, mode, par, reft, refidx, z, U)   # only to learn the consumption pattern of the uniforms
            r = H.r_csmc_run(mode, reft, T, N, z, H.csmc_uniform_queue(mode, U, o, N))
            out[f"mode{k}"] = np.array(mode); out[f"ref{k}"] = reft; out[f"z{k}"] = z; out[f"U{k}"] = U
            for key in ("logNC", "ancestors", "values", "logw", "resampled", "refidx", "lines", "line_values"):
                out[f"{key}{k}"] = np.asarray(r[key])
            k += 1
    out["ncases"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "csmc_kat.npz"), **out)


EXTRA = {"blockpf": blockpf_fixture, "csmc": csmc_fixture}


def main():
    assert H.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    if len(sys.argv) > 1:   # regenerate selected fixtures only
        for name in sys.argv[1:]:
            EXTRA[name]()
        return
    rng = np.random.default_rng(20251019)

    # ---- 1. stableLogSumWeights / GetESS known answers (population.h:46-50, sampler.h:413-417)
    out = {}
    cases = [-np.sqrt(np.arange(1.0, 5.0))]  # skeleton example, inst/skeleton/rcppsmc_hello_world.Rd:29-32
    for n in (1, 2, 3, 17, 256, 1000):
        cases.append(4.0 * rng.standard_normal(n) - 300.0)
    for i, lw in enumerate(cases):
        out[f"lw{i}"] = lw
        out[f"lse{i}"] = np.array(H.r_lse(lw))
        out[f"ess{i}"] = np.array(H.r_ess(lw))
    out["ncases"] = np.array(len(cases))
    np.savez(os.path.join(HERE, "lse_kat.npz"), **out)

    # ---- 2. sampler::Resample known answers (sampler.h:779-868)
    out = {}
    k = 0
    # Appendix B of SURVEY.md (8 particles)
    w8 = np.array([0.05, 0.30, 0.02, 0.20, 0.01, 0.25, 0.10, 0.07])
    appb = [(H.SYSTEMATIC, [0.37], None), (H.STRATIFIED, [0.37, 0.91, 0.05, 0.50, 0.66, 0.12, 0.99, 0.25, 0.40], None),
            (H.MULTINOMIAL, None, [0, 3, 0, 2, 0, 2, 1, 0])]
    for mode, U, counts in appb:
        idx, val = H.r_resample(mode, np.log(w8) + 3.0, uniforms=U, multinomial=counts)
        out[f"mode{k}"] = np.array(mode); out[f"lw{k}"] = np.log(w8) + 3.0
        out[f"U{k}"] = np.array(U if U is not None else []); out[f"counts{k}"] = np.array(counts if counts is not None else [], dtype=np.int32)
        out[f"w{k}"] = np.array([]); out[f"idx{k}"] = idx
        k += 1
    for n in (1, 2, 5, 33, 257, 1000, 4097):
        for mode in (H.SYSTEMATIC, H.STRATIFIED):
            for spread in (1.0, 6.0):
                lw = spread * rng.standard_normal(n)
                U = rng.random(1 if mode == H.SYSTEMATIC else n + 1)
                # (a) the reference's own path: weights = exp(lw - LSE)
                idx, _ = H.r_resample(mode, lw, uniforms=U)
                out[f"mode{k}"] = np.array(mode); out[f"lw{k}"] = lw; out[f"U{k}"] = U
                out[f"counts{k}"] = np.array([], dtype=np.int32); out[f"w{k}"] = np.array([]); out[f"idx{k}"] = idx
                k += 1
                # (b) identical normalised weights on the 2^-52 grid fed to the reference's cumsum
                w = H.random_weights(rng, n, spread)
                idx, _ = H.r_resample(mode, lw, uniforms=U, weights_override=w)
                out[f"mode{k}"] = np.array(mode); out[f"lw{k}"] = lw; out[f"U{k}"] = U
                out[f"counts{k}"] = np.array([], dtype=np.int32); out[f"w{k}"] = w; out[f"idx{k}"] = idx
                k += 1
        # injected multinomial counts -> index map (sampler.h:846-857)
        for total in (n, max(0, n - 1)):
            counts = rng.multinomial(total, rng.dirichlet(np.full(n, 0.4))).astype(np.int32)
            idx, _ = H.r_resample(H.MULTINOMIAL, np.zeros(n), multinomial=counts)
            out[f"mode{k}"] = np.array(H.MULTINOMIAL); out[f"lw{k}"] = np.zeros(n); out[f"U{k}"] = np.array([])
            out[f"counts{k}"] = counts; out[f"w{k}"] = np.array([]); out[f"idx{k}"] = idx
            k += 1
    out["ncases"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "resample_kat.npz"), **out)

    # ---- 3. pfNonlinBS traces from the reference sampler + nonlinbs_move on injected draws
    out = {}
    k = 0
    for (N, T, mode, thr) in ((200, 20, H.SYSTEMATIC, 1.01 * 200), (200, 20, H.STRATIFIED, 1.01 * 200), (64, 12, H.SYSTEMATIC, 0.5),
                              (1, 4, H.SYSTEMATIC, 1.01), (3, 6, H.STRATIFIED, 1.01 * 3)):
        _, y = H.sim_nonlin(T, 100 + k)
        z = rng.standard_normal(N * T)
        U = rng.random(T * (N + 1))
        r = H.r_pfnonlinbs(y, N, mode, thr, z, U)
        out[f"N{k}"] = np.array(N); out[f"T{k}"] = np.array(T); out[f"mode{k}"] = np.array(mode); out[f"thr{k}"] = np.array(thr)
        out[f"y{k}"] = y; out[f"z{k}"] = z; out[f"U{k}"] = U
        for key in ("mean", "sd", "logNC", "ESS", "resampled"):
            out[f"{key}{k}"] = r[key]
        k += 1
    out["ncases"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "pfnonlinbs_kat.npz"), **out)
    # ---- 4. pfLineartBS traces from the reference sampler + pflineart_move on injected draws
    out = {}
    k = 0
    for (N, T, mode, thr) in ((150, 30, H.SYSTEMATIC, 0.5), (150, 30, H.STRATIFIED, 0.5), (40, 10, H.SYSTEMATIC, 41.0)):
        data = H.sim_lineart(T, 200 + k)
        z = rng.standard_normal(4 * N * T)
        U = rng.random(T * (N + 1))
        r = H.r_pflineart(data, N, mode, thr, z, U)
        out[f"N{k}"] = np.array(N); out[f"T{k}"] = np.array(T); out[f"mode{k}"] = np.array(mode); out[f"thr{k}"] = np.array(thr)
        out[f"data{k}"] = data; out[f"z{k}"] = z; out[f"U{k}"] = U
        for key in ("Xm", "Xv", "Ym", "Yv", "logNC", "ESS", "resampled"):
            out[f"{key}{k}"] = r[key]
        k += 1
    out["ncases"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "pflineart_kat.npz"), **out)
    # ---- 5. the radiata data set (42 x 3: y, x1, x2), bundled data of the reference (inst/rawdata/radiata-data.csv)
    ref_dir = os.environ.get("SMC_REF_DIR", "/root/reference")
    raw = np.loadtxt(os.path.join(ref_dir, "inst", "rawdata", "radiata-data.csv"), skiprows=1)
    assert raw.shape == (42, 3)
    np.savez(os.path.join(HERE, "radiata.npz"), data=raw)
    for f in EXTRA.values():
        f()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
