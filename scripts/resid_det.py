"""Scratch: run-to-run determinism of the RESIDUAL / MULTINOMIAL engine paths on ONE handle."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rcppsmc_b200 as smc
import helpers as H
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(n, T, seed=1); f.set_data(y)
seq = [smc.RESIDUAL, smc.RESIDUAL, smc.MULTINOMIAL, smc.MULTINOMIAL, smc.RESIDUAL, smc.SYSTEMATIC, smc.RESIDUAL, smc.RESIDUAL]
first = {}
for k, mode in enumerate(seq):
    f.run(mode); f.sync(); o = f.read()
    key = mode
    if key not in first:
        first[key] = (o["logNC"].copy(), o["ESS"].copy())
        print(k, mode, "first run", o["logNC"][-1])
    else:
        a = first[key][0]
        diff = np.flatnonzero(a != o["logNC"])
        print(k, mode, "same as first" if diff.size == 0 else f"differs from step {diff[0]} (ESS there {first[key][1][diff[0]]:.1f} vs {o['ESS'][diff[0]]:.1f})", o["logNC"][-1])
f.close()
