"""Gene sharding for one-process-per-GPU launches (torchrun): which columns a rank owns and how the engine of a
rank is created.  The exchange steps themselves (two allreduces, one broadcast) run inside libruviii_b200.so
over NCCL; torch.distributed is used only to hand the NCCL unique id to every rank.

Partitioning rule (identical to the single-process multi-GPU split in engine.cu::do_plan):
rank g owns genes [floor(n g / N), floor(n (g+1) / N)).
"""
from __future__ import annotations

import numpy as np


def gene_range(n, rank, world):
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_columns(a, rank, world):
    """Columns of a row-major (rows x genes) matrix, or entries of a per-gene vector, owned by `rank` (a view)."""
    a = np.asarray(a)
    g0, g1 = gene_range(a.shape[-1], rank, world)
    return a[..., g0:g1]


def assemble_columns(parts):
    """Inverse of shard_columns over all ranks, in rank order."""
    return np.concatenate(parts, axis=-1)


def engine_from_torch_distributed(device=None):
    """Create this rank's engine inside an initialised torch.distributed job (any backend)."""
    import torch.distributed as dist

    from ._cabi import Engine
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", rank))
    if world == 1:
        return Engine(devices=[device])
    ids = [Engine.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return Engine(rank=rank, world=world, device=device, nccl_id=ids[0])
