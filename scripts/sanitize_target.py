"""Small runs of every device path for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rcppsmc_b200 as smc
import helpers as H
T = 6
_, y = H.sim_nonlin(T, 1)
for mode in (smc.SYSTEMATIC, smc.STRATIFIED, smc.MULTINOMIAL, smc.RESIDUAL):
    r = smc.pfNonlinBS(y, 5000 + mode, resample_mode=mode, seed=3)
    assert np.isfinite(r["mean"]).all()
d = H.sim_lineart(8, 1)
smc.pfLineartBS(d, 3001, resample_mode=smc.RESIDUAL, threshold=0.5, seed=2)
smc.LinRegLA_adapt(H.radiata()[:, :2], 512, 0.5, 0.9, seed=5)
smc.nonLinPMMH(H.sim_nonlin(10, 7, var_init=5.0, cos_offset=0.0)[1], 700, 2, seed=9)
yb = H.sim_lg(12, 5, 20.0)
o = smc.blockpfGaussianOpt(yb, 1500, 3, seed=3, diagnostics=True)
assert o["resampled"].sum() >= 1
x, yl = H.sim_lgss(10, 4)
ref = x[:, None] + 0.3 * np.random.default_rng(1).standard_normal((10, 3))
smc.compareNCestimates(yl, 300, 3, (0.9, 1.0, 0.0, 1.0), ref, seed=7)
h = smc.conditionalBPF(yl, 333, (0.9, 1.0, 0.0, 1.0), [0, 1, 2, 3], np.tile(x, (4, 1)), runs_per_chain=2, history=True)
smc.ancestor_lines(h["ancestors"][0], h["values"][0])
print("sanitize target ok")
