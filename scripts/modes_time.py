"""Device-resident time of one pfNonlinBS filter per resampling mode (scratch; not the bench contract)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
N = 1 << logn
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(N, T, seed=1)
f.set_data(y)
for mode, name in ((smc.SYSTEMATIC, "systematic"), (smc.STRATIFIED, "stratified"), (smc.MULTINOMIAL, "multinomial"), (smc.RESIDUAL, "residual")):
    ms = []
    for it in range(3):
        f.run(mode); f.sync(); ms.append(f.elapsed_ms())
    out = f.read()
    print(f"N=2^{logn} T={T} {name:12s} best={min(ms[1:]):8.3f} ms  {N * T / min(ms[1:]) * 1e3:.3e} particle-steps/s  logNC={out['logNC'][-1]:.4f}", flush=True)
f.close()
