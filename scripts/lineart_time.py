"""Scratch: pfLineartBS end-to-end time per resampling mode (2^20 x 250)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H
T, N = 250, 1 << 20
data = H.sim_lineart(T, 1)
for mode, name in ((smc.STRATIFIED, "stratified"), (smc.SYSTEMATIC, "systematic"), (smc.MULTINOMIAL, "multinomial"), (smc.RESIDUAL, "residual")):
    smc.pfLineartBS(data, N, resample_mode=mode, threshold=0.5, seed=3)
    t0 = time.perf_counter(); r = smc.pfLineartBS(data, N, resample_mode=mode, threshold=0.5, seed=3, full=True) if False else smc.pfLineartBS(data, N, resample_mode=mode, threshold=0.5, seed=3); t = time.perf_counter() - t0
    print(f"{name:12s} {1e3 * t:8.2f} ms  {N * T / t:.3e} particle-steps/s", flush=True)
