"""Multi-GPU plumbing: one process per GPU (torchrun), variants sharded, N-vectors replicated.

Variants shard naturally: (1/M) sum_j g_j (g_j' b) is a sum of independent rank-1 terms (the reference
splits the same loop over TBB threads, saige_fitnull.cpp:479).  Rank r of R owns the contiguous block
[r*M/R, (r+1)*M/R); one sum all-reduce of the N x k product block per GRM product (NCCL over NVLink) is
the only data-path collective.  torch.distributed is used for the rendezvous only.
"""
from __future__ import annotations

import numpy as np


def shard_range(m_total: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the variants owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(m_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a small byte string over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def init_comm_from_torch(ctx) -> None:
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed carries it."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    uid = ctx.comm_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, 128, 0)
    ctx.comm_init(uid, dist.get_rank(), dist.get_world_size())


def allreduce_sum_numpy(x: np.ndarray) -> np.ndarray:
    """Host-side sum all-reduce (gloo tests of the sharding logic)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
