"""Per-kernel timing of the island mode: python -m torch.distributed.run --nproc-per-node G scripts/island_time.py [log2 n per gpu] [T]"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
import rcppsmc_b200 as smc
from rcppsmc_b200 import island

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
_, y = H.sim_nonlin(T, 1)
f = smc.PfNonlinBS(1 << logn, T, seed=1, island=(rank, world))
island.connect(f)
f.set_data(y); f.set_profile(True)
os.environ["SMCB200_PROF_PRINT"] = "1"
for it in range(3):
    dist.barrier()
    f.run(smc.SYSTEMATIC); f.sync()
mv, nm, sc, ns = f.kernel_ms()
print(f"rank {rank}/{world}: total={f.elapsed_ms():.3f} ms  k_step={1e3*mv/nm:.1f} us  mass+scan={1e3*sc/ns:.1f} us", flush=True)
dist.barrier()
f.close()
