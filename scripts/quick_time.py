"""Scratch timing of the pfNonlinBS filter (device-resident), printed per N.  Not the bench contract."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H

T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for logn in [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["16", "20", "24"])]:
    N = 1 << logn
    _, y = H.sim_nonlin(T, 1)
    f = smc.PfNonlinBS(N, T, seed=1)
    f.set_data(y)
    for mode in (smc.SYSTEMATIC, smc.STRATIFIED):
        ms = []
        for it in range(5):
            f.run(mode)
            f.sync()
            ms.append(f.elapsed_ms())
        best = min(ms[1:])
        out = f.read()
        print(f"N=2^{logn} T={T} mode={mode} ms={['%.3f' % m for m in ms]} best={best:.3f} "
              f"psteps/s={N * T / best * 1e3:.3e} alg GB/s={N * T * 56 / best / 1e6:.1f} logNC={out['logNC'][-1]:.6f} mean[-1]={out['mean'][-1]:.4f}",
              flush=True)
    f.close()
