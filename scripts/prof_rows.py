"""Profiling targets for the widened rows (ncu): python scripts/prof_rows.py lineart|blockpf|linreg|csmc|small|draws"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
import rcppsmc_b200 as smc
which = sys.argv[1]
if which == "lineart":
    smc.pfLineartBS(H.sim_lineart(40, 1), 1 << 20, resample_mode=smc.SYSTEMATIC, threshold=0.5, seed=3)
elif which == "blockpf":
    smc.blockpfGaussianOpt(H.sim_lg(30, 5, 10.0), 1 << 20, 5, seed=3)
elif which == "linreg":
    smc.LinRegLA_adapt(H.radiata()[:, :2], 1 << 18, 0.5, 0.9, seed=5)
elif which == "csmc":
    x, y = H.sim_lgss(50, 4)
    ref = x[:, None] + 0.3 * np.random.default_rng(1).standard_normal((50, 20))
    smc.compareNCestimates(y, 1000, 20, (0.9, 1.0, 0.0, 1.0), ref, seed=7)
elif which == "small":
    f = smc.PfNonlinBS(1 << 16, 200, seed=1); f.set_data(H.sim_nonlin(200, 1)[1]); f.run(smc.SYSTEMATIC); f.sync(); f.close()
elif which == "draws":
    f = smc.PfNonlinBS(1 << 24, 5, seed=1); f.set_data(H.sim_nonlin(5, 1)[1]); f.run(smc.MULTINOMIAL); f.sync(); f.close()
print("ok", which)
