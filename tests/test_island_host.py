"""CPU: host-side island plumbing on a 2-rank gloo group (handle all-gather, sharding arithmetic)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rcppsmc_b200 import island
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank + 1]) * 64
    got = island.gather_handles(mine)
    first, n_loc = island.shard(1 << 20, rank, world)
    q.put((rank, got, first, n_loc))
    dist.destroy_process_group()


def test_handle_allgather_and_sharding_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, first, n_loc in res:
        assert got == [bytes([1]) * 64, bytes([2]) * 64]          # ordered by rank, identical on both ranks
        assert (first, n_loc) == (rank * (1 << 19), 1 << 19)


def test_shard_rejects_uneven_or_odd_islands():
    sys.path.insert(0, ROOT)
    from rcppsmc_b200 import island
    with pytest.raises(ValueError):
        island.shard(1001, 0, 2)
    with pytest.raises(ValueError):
        island.shard(6, 0, 2)       # 3 per island: Philox pairs particles
    assert island.shard(1000, 3, 4) == (750, 250)


def _chain_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rcppsmc_b200 import chains
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # stand-in for one finished chain per rank (the GPU work itself is covered by the -m gpu tests)
    mine = {"samples_sigv": np.full(5, 10.0 + rank), "loglike": np.arange(5.0) * (rank + 1), "accept": np.array(0.25 * (rank + 1))}
    got = chains.gather_chains(mine)
    q.put((rank, None if got is None else {k: v.tolist() for k, v in got.items()}, chains.chain_seed(7, rank)))
    dist.destroy_process_group()


def test_independent_chains_gather_two_ranks():
    # independent PMMH chains / filters shard one per rank without any data-path collective; only results are gathered
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chain_worker, args=(r, 2, 29633, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, got0, s0), (r1, got1, s1) = res
    assert got1 is None and s0 != s1
    assert got0["samples_sigv"] == [[10.0] * 5, [11.0] * 5]
    assert got0["loglike"] == [[0.0, 1.0, 2.0, 3.0, 4.0], [0.0, 2.0, 4.0, 6.0, 8.0]]
    assert got0["accept"] == [0.25, 0.5]
