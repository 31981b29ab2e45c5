import ctypes, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ["SMCB200_LIB"] = "build/lib_lbprof.so"
import numpy as np
import rcppsmc_b200 as smc
import helpers as H
N, T = 1 << 24, 20
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(N, T, seed=1)
f.set_data(y); f.set_profile(True)
L = smc.lib()
st = (ctypes.c_ulonglong * 8)()
f.run(smc.SYSTEMATIC); f.sync()
L.smcb200_dbg_lookback_stats(st, 1)
f.run(smc.SYSTEMATIC); f.sync()
L.smcb200_dbg_lookback_stats(st, 1)
cyc, rounds, calls, stalled = st[0], st[1], st[2], st[3]
mv, nm, sc, ns = f.kernel_ms()
print(f"k_scan={1e3*sc/ns:.1f} us/launch; lookback calls={calls} ({calls/ (T)} per launch) avg cycles/call={cyc/calls:.0f} avg rounds={rounds/calls:.2f} stalled polls/call={stalled/calls:.2f}")
