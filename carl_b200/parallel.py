"""Sharding of the env batch over the GPUs of one box (one process per GPU).

The env instances are independent, so the path shards with no data-path collective except the
one the north star names: the gathered observation tensor. Two implementations:

* ``ObsGather(mode="nccl")`` -- ``torch.distributed.all_gather_into_tensor`` (baseline).
* ``ObsGather(mode="fused")`` -- the step / reset / rollout kernels store each obs row straight into
  every rank's gathered buffer over NVLink (peer-mapped or NVLS-multicast symmetric memory) and
  exchange per-rank push counts; no collective launch, no host synchronisation, CUDA-graph
  capturable. ``pipelined=False``: when a launch has completed, the gathered tensor of ITS
  observation is complete (the consumer that feeds obs k into step k+1). ``pipelined=True``: the
  gathered tensor of the PREVIOUS launch's observation is complete -- the transfer of obs k
  overlaps the physics of launch k+1 (``gather(lag=1)`` is then free, ``gather(lag=0)`` appends a
  flush launch).
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

    def __init__(self, env: Any, mode: str = "nccl", group: Any = None, pipelined: bool = False,
                 symmetric: str = "auto"):
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
        self._pending = None
        self._pending_buf = None
        self._pending_src = None
        self._alt = None
        self.pipelined = bool(pipelined)
        self.transport = "nccl"
        if mode == "fused":
            self._setup_fused(symmetric)
        elif mode != "nccl":
            raise ValueError(f"unknown gather mode {mode!r}")

    def gather(self, obs=None, lag: int = 0):
        """Returns the ``[N_global, D]`` tensor of the latest observation-producing call (``lag=0``)
        or of the one before it (``lag=1``: a pipelined consumer enqueues launch k+1 first and then
        asks for launch k, so neither the fused wait nor the NCCL collective stalls the stream).

        nccl: collective on the current stream. fused: the rows were already stored into every
        rank's buffer by the step/reset/rollout kernel itself; only a one-warp wait kernel is enqueued
        (no NCCL call, no host sync)."""
        if self.mode == "fused":
            import ctypes

            import torch

            from carl_b200 import _native

            ptr = ctypes.c_void_p()
            _native.check(self._lib.carlb_gather_wait(self._g, int(lag), torch.cuda.current_stream(self.env.device).cuda_stream,
                                                      ctypes.byref(ptr)))
            return self._views[ptr.value]
        obs = self.env._obs if obs is None else obs
        if lag == 1 and self.equal:
            # pipelined NCCL: start this launch's collective asynchronously into the other buffer and
            # hand back the previous one (completed meanwhile)
            import torch

            if self._pending is None:
                self._alt = torch.zeros_like(self.gathered)
            prev, prev_buf = self._pending, self._pending_buf
            buf = self._alt if self._pending_buf is self.gathered else self.gathered
            src = obs.clone()  # the env's obs buffer is overwritten by the next launch
            self._pending = self.dist.all_gather_into_tensor(buf, src, group=self.group, async_op=True)
            self._pending_buf, self._pending_src = buf, src
            if prev is None:
                self._pending.wait()
                return buf
            prev.wait()
            return prev_buf
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

    def _setup_fused(self, symmetric: str = "auto"):
        """Create this rank's symmetric buffer, map every peer's buffer and attach the gather to the env
        handle (its kernels then push obs rows to all ranks over NVLink and publish push counts).

        ``symmetric``: "torch" -- the block comes from ``torch.distributed._symmetric_memory`` (plumbing:
        allocation + rendezvous), which also yields the NVLS multicast alias, so one ``multimem.st`` per
        row reaches every rank; "ipc" -- a cudaMalloc block shared through CUDA IPC handles (one store
        per rank and row); "auto" -- torch symmetric memory when every rank gets it, else IPC."""
        import ctypes
        import os

        import torch

        from carl_b200 import _native

        env = self.env
        self._lib = _native.load()
        self._g = ctypes.c_void_p()
        self._symm = None
        D = env._info.obs_dim
        world = env.world_size
        symmetric = os.environ.get("CARLB_GATHER_SYMMETRIC", symmetric)
        use_symm = symmetric in ("auto", "torch") and world > 1
        if use_symm:
            ok, err = 1, None
            try:
                import torch.distributed._symmetric_memory as symm_mem

                nbytes = int(self._lib.carlb_gather_bytes(env.global_num_envs, D))
                block = symm_mem.empty(nbytes, dtype=torch.uint8, device=env.device)
                block.zero_()
                torch.cuda.synchronize(env.device)
                hdl = symm_mem.rendezvous(block, self.group if self.group is not None else self.dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                mc = int(hdl.multicast_ptr)  # 0 when the fabric has no NVLS multicast
                if os.environ.get("CARLB_GATHER_MULTICAST", "1") == "0":
                    mc = 0
                if len(ptrs) != world or any(p == 0 for p in ptrs):
                    raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
            except Exception as e:  # pragma: no cover - depends on the box / driver
                ok, err = 0, e
            flag = torch.tensor([ok], device=env.device, dtype=torch.int32)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 1:
                arr = (ctypes.c_void_p * world)(*ptrs)
                _native.check(self._lib.carlb_gather_create_symmetric(env.device.index, env.rank, world, env.global_num_envs,
                                                                      D, arr, ctypes.c_void_p(mc or None), ctypes.byref(self._g)))
                self._symm = (block, hdl)  # keep the mapping alive
                self.transport = "symmetric-memory multicast (multimem.st)" if mc else "symmetric-memory peer stores"
            else:
                if symmetric == "torch":
                    raise RuntimeError(f"torch symmetric memory unavailable on some rank ({err})")
                use_symm = False
        if not use_symm:
            _native.check(self._lib.carlb_gather_create(env.device.index, env.rank, world, env.global_num_envs, D,
                                                        ctypes.byref(self._g)))
            handle = (ctypes.c_ubyte * 64)()
            _native.check(self._lib.carlb_gather_export(self._g, handle))
            mine = bytes(handle)
            if world > 1:
                handles = [None] * world
                self.dist.all_gather_object(handles, mine, group=self.group)
            else:
                handles = [mine]
            for r, h in enumerate(handles):
                if r == env.rank:
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                _native.check(self._lib.carlb_gather_open(self._g, r, buf))
            self.transport = "cuda-ipc peer stores"
        _native.check(self._lib.carlb_gather_attach(self._g, env._handle))
        _native.check(self._lib.carlb_gather_set_mode(self._g, 1 if self.pipelined else 0))
        slot_floats = (env.global_num_envs * D + 63) // 64 * 64
        torch.cuda.synchronize(env.device)
        # tensor views of the slots of the local buffer, created lazily from the pointers carlb_gather_wait returns
        self._views = _SlotViews(env.device, env.global_num_envs, D, slot_floats)
        if world > 1:
            self.dist.barrier(group=self.group)

    def resync(self):
        """After replaying CUDA graphs that contain obs-producing launches: re-read the device-side push
        counter (synchronises the current stream) so that ``gather()`` returns the right slot again."""
        if self.mode == "fused":
            import torch

            from carl_b200 import _native

            _native.check(self._lib.carlb_gather_resync(self._g, torch.cuda.current_stream(self.env.device).cuda_stream))

    def close(self):
        if self.mode == "fused" and getattr(self, "_g", None) is not None and self._g.value:
            self._lib.carlb_gather_destroy(self._g)
            import ctypes

            self._g = ctypes.c_void_p()
            self._symm = None


class _DevArray:
    def __init__(self, ptr: int, shape: tuple):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}


class _SlotViews(dict):
    """pointer -> torch view of a gathered slot (created on first use)."""

    def __init__(self, device, n, d, slot_floats):
        super().__init__()
        self.device, self.n, self.d = device, n, d

    def __missing__(self, ptr):
        import torch

        t = torch.as_tensor(_DevArray(ptr, (self.n, self.d)), device=self.device)
        self[ptr] = t
        return t


def host_gather_reference(shards: list[np.ndarray]) -> np.ndarray:
    """What the gathered tensor must equal: the shards concatenated in rank order."""
    return np.concatenate(shards, axis=0)
