"""pfNonlinBS size sweep (BASELINE configs[4]): device-resident time of one 100-step filter per population size, SYSTEMATIC."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H
T = 100
_, y = H.sim_nonlin(T, 1)
for logn in [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16,18,20,22,24,26".split(","))]:
    N = 1 << logn
    f = smc.PfNonlinBS(N, T, seed=1)
    f.set_data(y)
    ms = []
    for it in range(4):
        f.run(smc.SYSTEMATIC); f.sync(); ms.append(f.elapsed_ms())
    best = min(ms[1:])
    out = f.read()
    print(f"N=2^{logn:<2d} T={T}  {best:9.3f} ms  {N * T / best * 1e3:.3e} particle-steps/s  {100 * N * T * 56 / best / 1e6 / 6458.4:5.1f} % of HBM roofline  logNC={out['logNC'][-1]:.4f}", flush=True)
    f.close()
