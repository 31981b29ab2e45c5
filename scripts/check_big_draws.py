"""Bit-exactness of the native MULTINOMIAL / RESIDUAL definitions against the oracle at large N (guide-table search, residual shift)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, helpers as H, rcppsmc_b200 as smc
for logn in (22, 24, 25):
    n = (1 << logn) + (3 if logn == 25 else 0)
    rng = np.random.default_rng(logn)
    for spread in (2.0, 9.0):
        w = H.random_weights(rng, n, spread)
        U = rng.random(n)
        for mode in (H.MULTINOMIAL, H.RESIDUAL):
            idx, cnt = smc.resample(mode, w, U, return_counts=True)
            oidx, ocnt = H.o_resample_w(mode, w, U)
            bad = np.flatnonzero(cnt != ocnt)
            print(f"n={n} spread={spread} mode={mode} counts_equal={bad.size == 0} idx_equal={np.array_equal(idx, oidx)} max_count={cnt.max()} sum={cnt.sum()}", flush=True)
