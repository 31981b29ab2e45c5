"""Scratch: BASELINE configs[0] (pfNonlinBS, 1000 particles x 50 steps) per resampling scheme, one-shot C-ABI call with host buffers."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
import rcppsmc_b200 as smc
T, N = 50, 1000
_, y = H.sim_nonlin(T, 1)
for mode, name in ((smc.MULTINOMIAL, "multinomial (reference default)"), (smc.SYSTEMATIC, "systematic")):
    smc.pfNonlinBS(y, N, resample_mode=mode, seed=1)
    ts = []
    for r in range(20):
        t0 = time.perf_counter(); smc.pfNonlinBS(y, N, resample_mode=mode, seed=2 + r); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    print(f"pfNonlinBS 1000 x 50 {name}: {1e6 * t:.0f} us per filter end to end = {N * T / t:.3e} particle-steps/s", flush=True)
if H.ref_available():
    import ctypes
    R = H.ref()
    R.smcref_time_pfnonlinbs.restype = ctypes.c_double
    yy = np.ascontiguousarray(y)
    sec = min(R.smcref_time_pfnonlinbs(yy.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_long(T), ctypes.c_long(N), 3, ctypes.c_double(1.01 * N)) for _ in range(5))
    print(f"reference CPU path (1 core, systematic): {1e6 * sec:.0f} us per filter = {N * T / sec:.3e} particle-steps/s")
