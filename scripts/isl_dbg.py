import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
import rcppsmc_b200 as smc
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = 1 << logn
_, y = H.sim_nonlin(T, 1)
def run(force, dbg):
    os.environ["SMCB200_FORCE_ISLAND"] = "1" if force else "0"
    os.environ["SMCB200_DBG_STEP"] = str(dbg)
    os.environ["SMCB200_SMALL_MAX"] = "0"
    f = smc.PfNonlinBS(N, T, seed=1); f.set_data(y); f.run(smc.SYSTEMATIC); f.sync()
    out = f.read(); anc = f.read_ancestors(); f.close()
    return out, anc
ref, anc0 = run(False, 0)
for dbg in (0, 256):
    o, a = run(True, dbg)
    print("dbg", dbg, "logNC", o["logNC"], "ref", ref["logNC"], "ESS diff", np.max(np.abs(o["ESS"] - ref["ESS"])), "anc equal", np.array_equal(a, anc0), flush=True)
