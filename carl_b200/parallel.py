"""Sharding of the env batch over the GPUs of one box (one process per GPU).

The env instances are independent, so the path shards with no data-path collective except the
one the north star names: the gathered observation tensor. Two implementations:

* ``ObsGather(mode="nccl")`` -- ``torch.distributed.all_gather_into_tensor`` (baseline).
* ``ObsGather(mode="fused")`` -- the step kernel stores each obs row straight into every rank's
  gathered buffer through P2P-mapped memory (CUDA IPC handles exchanged once); a step then only
  needs a cross-rank barrier, not a separate collective launch.
"""
from __future__ import annotations

from typing import Any

import numpy as np


def shard_range(n_global: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced partition: rank r owns ``[lo, hi)``; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_global: int, world_size: int) -> list[int]:
    return [shard_range(n_global, r, world_size)[1] - shard_range(n_global, r, world_size)[0] for r in range(world_size)]


class ObsGather:
    """All-gather of the per-rank observation shards into ``[N_global, D]`` on every rank."""

    def __init__(self, env: Any, mode: str = "nccl", group: Any = None):
        import torch
        import torch.distributed as dist

        self.env = env
        self.mode = mode
        self.group = group
        self.dist = dist
        self.world_size = env.world_size
        self.sizes = shard_sizes(env.global_num_envs, env.world_size)
        D = env._info.obs_dim
        self.equal = len(set(self.sizes)) == 1
        self.gathered = torch.zeros(env.global_num_envs, D, dtype=torch.float32, device=env.device)
        self._padded = None
        self._mine = None
        if mode == "fused":
            self._setup_fused()
        elif mode != "nccl":
            raise ValueError(f"unknown gather mode {mode!r}")

    def gather(self, obs=None):
        """NCCL path: collective on the current stream. Returns the ``[N_global, D]`` tensor."""
        obs = self.env._obs if obs is None else obs
        if self.mode == "fused":
            # rows were already stored by the step kernel; only order the ranks
            self.dist.barrier(group=self.group)
            return self.gathered
        if self.equal:
            self.dist.all_gather_into_tensor(self.gathered, obs, group=self.group)
        else:
            # ragged shards (sizes differ by one): gather max-size padded rows, then compact
            import torch

            m = max(self.sizes)
            if self._padded is None:
                self._padded = torch.zeros(self.world_size * m, obs.shape[1], dtype=obs.dtype, device=obs.device)
                self._mine = torch.zeros(m, obs.shape[1], dtype=obs.dtype, device=obs.device)
            self._mine[: obs.shape[0]].copy_(obs)
            self.dist.all_gather_into_tensor(self._padded, self._mine, group=self.group)
            off = 0
            for r, sz in enumerate(self.sizes):
                self.gathered[off:off + sz].copy_(self._padded[r * m:r * m + sz])
                off += sz
        return self.gathered

    def _setup_fused(self):
        raise NotImplementedError("fused peer-store gather is enabled in carl_b200.fused_gather")


def host_gather_reference(shards: list[np.ndarray]) -> np.ndarray:
    """What the gathered tensor must equal: the shards concatenated in rank order."""
    return np.concatenate(shards, axis=0)
