"""Env-batch sharding across ranks (one process per GPU) and the single data-path collective.

Environments are independent, so the global batch is split into contiguous ranges, one per rank;
actions are sliced, and the only exchange is one all-gather of the packed per-step outputs
(north_star: "at most one NCCL all-gather ... for the returned observation tensor").
Backend-agnostic (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of the envs owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_outputs(obs, reward, discount, step_type):
    """[n, obs_dim + 3] float64 block so that one collective moves everything."""
    import torch
    return torch.cat([obs, reward[:, None], discount[:, None], step_type.to(obs.dtype)[:, None]], dim=1)


def all_gather_outputs(local_block, world: int, n_total: int):
    """All-gather the packed outputs of equally sized shards -> [n_total, obs_dim + 3].

    `all_gather_into_tensor` needs the same shard size on every rank: the global batch must divide evenly
    (shard_range then gives every rank n_total / world envs)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_block
    if n_total % world != 0 or local_block.shape[0] * world != n_total:
        raise ValueError(f"all_gather_outputs needs equal shards: n_total={n_total}, world={world}, "
                         f"local rows={local_block.shape[0]}")
    out = torch.empty((world * local_block.shape[0], local_block.shape[1]), dtype=local_block.dtype,
                      device=local_block.device)
    dist.all_gather_into_tensor(out, local_block.contiguous())
    return out[:n_total]
