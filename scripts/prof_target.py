"""Short profiling target for ncu: one pfNonlinBS filter, 2^24 particles x T steps, systematic resampling."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import rcppsmc_b200 as smc
import helpers as H

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = int(sys.argv[3]) if len(sys.argv) > 3 else smc.SYSTEMATIC
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(1 << logn, T, seed=1)
f.set_data(y)
f.run(mode)
f.sync()
print("ok", f.elapsed_ms(), f.read()["logNC"][-1])
f.close()
