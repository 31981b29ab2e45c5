import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rcppsmc_b200 as smc
import helpers as H
T, N, S = 50, 1000, 100
x, yy = H.sim_lgss(T, 4)
ref = x[:, None] + 0.3 * np.random.default_rng(1).standard_normal((T, S))
for ind in (False, True):
    smc.compareNCestimates(yy, N, S, (0.9, 1.0, 0.0, 1.0), ref, seed=7, independent_sims=ind)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); a = smc.compareNCestimates(yy, N, S, (0.9, 1.0, 0.0, 1.0), ref, seed=7, independent_sims=ind); ts.append(time.perf_counter() - t0)
    print("compareNC independent", ind, min(ts), "s; means", a["smcOut"].mean(axis=0), a["csmcOut"].mean(axis=0))
