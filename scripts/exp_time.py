"""Scratch experiment driver: per-kernel times of one pfNonlinBS filter (env SMCB200_DBG selects experiments)."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = 1 << logn
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(N, T, seed=1)
f.set_data(y); f.set_profile(True)
for it in range(3):
    f.run(smc.SYSTEMATIC); f.sync()
mv, nm, sc, ns = f.kernel_ms()
print(f"STEP={os.environ.get('SMCB200_DBG_STEP','0')} SCAN={os.environ.get('SMCB200_DBG_SCAN','0')} N=2^{logn} total={f.elapsed_ms():.3f} ms  k_step={1e3*mv/nm:.1f} us/launch  k_scan={1e3*sc/ns:.1f} us/launch", flush=True)
f.close()
