import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, helpers as H, rcppsmc_b200 as smc
for logn in (22, 24, 26):
    n = 1 << logn
    rng = np.random.default_rng(logn)
    w = H.random_weights(rng, n, 2.0)
    for mode in (H.SYSTEMATIC, H.STRATIFIED):
        U = rng.random(1 if mode == H.SYSTEMATIC else n + 1)
        idx, cnt = smc.resample(mode, w, U, return_counts=True)
        oidx, ocnt = H.o_resample_w(mode, w, U)
        bad = np.flatnonzero(cnt != ocnt)
        print(f"n=2^{logn} mode={mode} counts_equal={bad.size == 0} idx_equal={np.array_equal(idx, oidx)} first_bad={bad[:5]} sum={cnt.sum()} osum={ocnt.sum()}", flush=True)
