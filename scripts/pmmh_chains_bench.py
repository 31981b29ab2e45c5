"""Independent PMMH chains, one per GPU (BASELINE configs[3]): python -m torch.distributed.run --nproc-per-node G scripts/pmmh_chains_bench.py
Prints the aggregate filter rate (all chains) measured between two barriers; no data-path collective."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
import rcppsmc_b200 as smc
from rcppsmc_b200 import chains

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
T, N, M = 500, 1 << 16, int(sys.argv[1]) if len(sys.argv) > 1 else 200
y = H.sim_nonlin(T, 7, var_init=5.0, cos_offset=0.0)[1]
smc.nonLinPMMH(y, N, 3, seed=1)   # warm-up (context, allocations)
dist.barrier()
t0 = time.perf_counter()
res = chains.pmmh_chains(y, N, M, base_seed=11)
dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    acc = res["MH_acceptance_rate"][:, -1]
    print(f"{world} independent chains x {M} iterations of 2^16-particle filters x {T} obs: {dt:.3f} s, {world * (M + 1) / dt:.1f} filters/s aggregate, "
          f"{world * (M + 1) * N * T / dt:.3e} particle-steps/s; posterior mean sigv per chain {np.round(res['samples_sigv'][:, M // 2:].mean(axis=1), 2)}; acceptance {np.round(acc, 2)}", flush=True)
dist.destroy_process_group()
