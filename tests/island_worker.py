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
    y_skew = y.copy()
    y_skew[T // 3] = 55.0      # an outlier: only the few particles near |x| = 33 keep any weight -> a handful of parents own
    y_skew[2 * T // 3] = -9.0  # everybody's children, most of them on another island (the surplus lists cross NVLink)
    ok = True
    for mode, data, thr in ((smc.SYSTEMATIC, y, None), (smc.STRATIFIED, y, None), (smc.SYSTEMATIC, y_skew, None), (smc.SYSTEMATIC, y_skew, 0.5)):
        f = smc.PfNonlinBS(n_loc, T, seed=77, island=(rank, world))
        exchange_handles(f)
        f.set_data(data)
        for rep in range(2):   # twice: mailbox sequence numbers must stay unique across runs
            f.run(mode, thr)
            f.sync()
        out = f.read()
        anc_loc = torch.from_numpy(f.read_ancestors().astype(np.int64))   # global parent ids of this island's slots (last step)
        anc_all = [torch.empty_like(anc_loc) for _ in range(world)]
        dist.all_gather(anc_all, anc_loc)
        dist.barrier()
        # every rank must hold identical traces
        mine = torch.from_numpy(np.concatenate([out["logNC"], out["ESS"], out["mean"], out["sd"]]))
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        same = all(torch.equal(allv[0], v) for v in allv)
        if rank == 0:
            g = smc.PfNonlinBS(n_glob, T, seed=77)
            g.set_data(data)
            g.run(mode, thr)
            ref = g.read()
            ref_anc = g.read_ancestors().astype(np.int64)
            g.close()
            anc_same = bool(np.array_equal(torch.cat(anc_all).numpy(), ref_anc))   # last step's ancestor indices, bit for bit
            # same Philox draws (global particle ids) and the same algorithm: only reduction order / rare boundary flips differ
            err = {k: float(np.max(np.abs(out[k] - ref[k]) / np.maximum(1e-12, np.abs(ref[k])))) for k in ("logNC", "ESS", "mean", "sd")}
            good = same and anc_same and np.array_equal(out["resampled"], ref["resampled"]) and err["logNC"] < 1e-9 and err["ESS"] < 1e-6 \
                and err["mean"] < 1e-6 and err["sd"] < 1e-6
            print(f"island world={world} N=2^{log2n} T={T} mode={mode} skew={data is y_skew} thr={thr}: identical_across_ranks={same} ancestors_bit_exact={anc_same} min_ESS/N={float(np.min(ref['ESS'])) / n_glob:.2e} rel_err_vs_1gpu={err} -> {'OK' if good else 'FAIL'}",
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
