"""Host-side plumbing of island mode (one process per GPU): sharding arithmetic and the all-gather of the 64-byte CUDA
IPC handles over an existing torch.distributed process group (any backend).  The data path itself never goes through
torch.distributed: after ``connect`` the kernels talk over NVLink peer memory."""
import numpy as np


def shard(n_global, rank, world):
    """(first global slot, local particle count) of ``rank``; islands are equal and even-sized (Philox pairs particles)."""
    if n_global % world:
        raise ValueError("the population must split evenly across ranks")
    n_loc = n_global // world
    if world > 1 and n_loc % 2:
        raise ValueError("island size must be even")
    return rank * n_loc, n_loc


def gather_handles(handle, group=None, device=None):
    """All-gathers one 64-byte handle per rank; returns the list ordered by rank."""
    import torch
    import torch.distributed as dist
    if len(handle) != 64:
        raise ValueError("a CUDA IPC handle is 64 bytes")
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.frombuffer(handle, dtype=np.uint8).copy())
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def connect(filter_handle, group=None, device=None):
    """Exchange IPC handles of an island-mode ``PfNonlinBS`` and map the peers' slabs."""
    filter_handle.connect(gather_handles(filter_handle.ipc_handle(), group, device))
