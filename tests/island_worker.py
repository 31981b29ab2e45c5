"""Worker of the island-mode test: run with `python -m torch.distributed.run --nproc-per-node G tests/island_worker.py`.
Every rank filters its island of ONE population with global resampling (peer-memory exchange inside the kernels); rank 0
also runs the same population on a single GPU and compares the traces."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import rcppsmc_b200 as smc  # noqa: E402


from rcppsmc_b200 import island  # noqa: E402


def exchange_handles(f):
    """all-gather of the 64-byte IPC handles over the host-side process group (gloo: plumbing, not data path)"""
    island.connect(f)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    n_glob = 1 << log2n
    n_loc = n_glob // world
    _, y = H.sim_nonlin(T, 3)
    ok = True
    for mode in (smc.SYSTEMATIC, smc.STRATIFIED):
        f = smc.PfNonlinBS(n_loc, T, seed=77, island=(rank, world))
        exchange_handles(f)
        f.set_data(y)
        for rep in range(2):   # twice: mailbox sequence numbers must stay unique across runs
            f.run(mode)
            f.sync()
        out = f.read()
        dist.barrier()
        # every rank must hold identical traces
        mine = torch.from_numpy(np.concatenate([out["logNC"], out["ESS"], out["mean"], out["sd"]]))
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        same = all(torch.equal(allv[0], v) for v in allv)
        if rank == 0:
            g = smc.PfNonlinBS(n_glob, T, seed=77)
            g.set_data(y)
            g.run(mode)
            ref = g.read()
            g.close()
            # same Philox draws (global particle ids) and the same algorithm: only reduction order / rare boundary flips differ
            err = {k: float(np.max(np.abs(out[k] - ref[k]) / np.maximum(1e-12, np.abs(ref[k])))) for k in ("logNC", "ESS", "mean", "sd")}
            good = same and np.array_equal(out["resampled"], ref["resampled"]) and err["logNC"] < 1e-9 and err["ESS"] < 1e-6 \
                and err["mean"] < 1e-6 and err["sd"] < 1e-6
            print(f"island world={world} N=2^{log2n} T={T} mode={mode}: identical_across_ranks={same} rel_err_vs_1gpu={err} -> {'OK' if good else 'FAIL'}",
                  flush=True)
            if not good:
                print("  per-step logNC diff", out["logNC"] - ref["logNC"], "\n  per-step mean diff", out["mean"] - ref["mean"], flush=True)
            ok = ok and good
        f.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
