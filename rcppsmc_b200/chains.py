"""Independent chains, one per GPU: PMMH chains or whole particle filters that share nothing.
Every rank runs its own chain on its own device; the only communication is the host-side gather of the finished
results over an existing torch.distributed process group (any backend).  No collective sits on the data path."""
import numpy as np


def chain_seed(base_seed, rank):
    """Distinct Philox keys per chain: the engine keys every draw by (seed, stream, step, particle), so chains with
    different seeds are independent streams."""
    return (int(base_seed) + 0x9E3779B97F4A7C15 * (int(rank) + 1)) & 0xFFFFFFFFFFFFFFFF


def gather_chains(result, group=None, dst=0):
    """Gathers one dict of equally shaped float arrays per rank on ``dst``; returns {key: array [world, ...]} there, None elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    keys = sorted(result)
    flat = np.concatenate([np.asarray(result[k], dtype=np.float64).reshape(-1) for k in keys])
    mine = torch.from_numpy(flat.copy())
    out = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, out, dst=dst, group=group)
    if rank != dst:
        return None
    stacked = np.stack([t.numpy() for t in out])
    res, off = {}, 0
    for k in keys:
        shape = np.asarray(result[k]).shape
        size = int(np.prod(shape)) if shape else 1
        res[k] = stacked[:, off:off + size].reshape((world,) + tuple(shape))
        off += size
    return res


def pmmh_chains(data, particles, iterations, base_seed=1, group=None, **kw):
    """One ``nonLinPMMH`` chain per rank (reference src/nonLinPMMH.cpp run once per chain); rank 0 gets the stacked chains."""
    import torch.distributed as dist
    from . import nonLinPMMH
    rank = dist.get_rank(group)
    out = nonLinPMMH(data, particles, iterations, seed=chain_seed(base_seed, rank), **kw)
    return gather_chains(out, group)
