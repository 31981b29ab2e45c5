"""Scratch: per-filter cost of nonLinPMMH (2^16-particle filters x 500 observations) from the slope between two chain lengths."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rcppsmc_b200 as smc
import helpers as H
T, N = 500, 1 << 16
y = H.sim_nonlin(T, 7, var_init=5.0, cos_offset=0.0)[1]
smc.nonLinPMMH(y, N, 3, seed=9)
ts = {}
for M in (20, 120):
    t0 = time.perf_counter(); r = smc.nonLinPMMH(y, N, M, seed=9); ts[M] = time.perf_counter() - t0
    print(f"M={M}: {ts[M]:.3f} s  acceptance {r['MH_acceptance_rate'][-1]:.2f}")
per = (ts[120] - ts[20]) / 100
print(f"per filter {1e3 * per:.2f} ms -> {1 / per:.1f} filters/s asymptotically; fixed cost {ts[20] - 21 * per:.3f} s")
