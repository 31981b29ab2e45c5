"""Scratch: nonLinPMMH at BASELINE configs[3] (2^16-particle filters x 500 observations), filters per second."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
import rcppsmc_b200 as smc
T, N = 500, 1 << 16
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100
y = H.sim_nonlin(T, 7, var_init=5.0, cos_offset=0.0)[1]
smc.nonLinPMMH(y, N, 4, seed=9)
t0 = time.perf_counter(); r = smc.nonLinPMMH(y, N, M, seed=9); t = time.perf_counter() - t0
print(f"nonLinPMMH 2^16 x 500, {M + 1} filters: {t:.3f} s = {(M + 1) / t:.1f} filters/s, accept rate {r['accept_rate'][-1] if 'accept_rate' in r else float('nan'):.3f}", flush=True)
