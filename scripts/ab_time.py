"""A/B timing of the pfNonlinBS headline filter: total device time and the per-kernel split, for the library named by
SMCB200_LIB (default: the in-tree build) and the scan implementation named by SMCB200_SCAN_IMPL.  Scratch tool."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rcppsmc_b200 as smc
import helpers as H

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mode = int(sys.argv[3]) if len(sys.argv) > 3 else smc.SYSTEMATIC
N = 1 << logn
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(N, T, seed=1)
f.set_data(y)
ms = []
for it in range(4):
    f.run(mode, 1.01 * N); f.sync(); ms.append(f.elapsed_ms())
out = f.read()
f.set_profile(True)
f.run(mode, 1.01 * N); f.sync()
mv, nm, sc, ns = f.kernel_ms()
best = min(ms[1:])
print(f"lib={os.path.basename(os.environ.get('SMCB200_LIB', 'default'))} impl={os.environ.get('SMCB200_SCAN_IMPL', '1')} N=2^{logn} T={T} mode={mode} "
      f"best={best:.3f} ms ({100 * N * T * 56 / best / 1e6 / 6458.4:.1f} % roofline) k_step={1e3 * mv / max(nm, 1):.1f} us scan={1e3 * sc / max(ns, 1):.1f} us "
      f"logNC={out['logNC'][-1]:.6f} mean[-1]={out['mean'][-1]:.5f} ess[-1]={out['ESS'][-1]:.1f}", flush=True)
f.close()
