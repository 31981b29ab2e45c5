import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import helpers as H
import rcppsmc_b200 as smc
T, N, lag = 100, 1 << 20, 5
y = H.sim_lg(T, 5, 10.0)
buf = torch.empty((T, N), dtype=torch.float64, device="cuda")
for k in range(4):
    t0 = time.perf_counter(); r = smc.blockpfGaussianOpt_device(y, N, lag, buf.data_ptr(), seed=3); t = time.perf_counter() - t0
    print(f"device-resident call {k}: {1e3*t:.1f} ms  logNC {r['logNC']:.6f}", flush=True)
