import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, helpers as H, rcppsmc_b200 as smc
data = H.radiata()[:, [0, 1]]
N = 1000
t0 = time.time()
try:
    out = smc.LinRegLA_adapt(data, N, 0.5, 0.9, seed=3)
    print("L", out["Temps"].size, "temps", out["Temps"][:6], "reps", out["mcmcRepeats"][:10], "logNC", out["logNC_standard"], out["logNC_ps_trap2"], f"{time.time()-t0:.1f}s")
except Exception as e:
    print("ERR", e)
