"""Mixed-env batch: several homogeneous shards (e.g. CARLPendulum + CARLAcrobot, BASELINE
config 3) advanced by ONE kernel launch (``carlb_mixed_step``)."""
from __future__ import annotations

import ctypes

import torch

from carl_b200 import _native
from carl_b200.envs.carl_env import _TORCH_ACT, CARLEnv


class MixedBatch:
    def __init__(self, envs: list[CARLEnv]):
        if not 1 <= len(envs) <= _native.MAX_MIXED:
            raise ValueError(f"between 1 and {_native.MAX_MIXED} shards per launch")
        if len({e.device for e in envs}) != 1:
            raise ValueError("all shards of a mixed batch must live on one device")
        self.envs = list(envs)
        self._lib = _native.load()
        self._handles = (ctypes.c_void_p * len(envs))(*[e._handle for e in envs])

    @property
    def num_envs(self) -> int:
        return sum(e.num_envs for e in self.envs)

    def reset(self, *, seed: int | None = None):
        outs = []
        for e in self.envs:
            outs.append(e.reset(seed=seed))
        return outs

    def step(self, actions: list[torch.Tensor]):
        assert len(actions) == len(self.envs)
        if not all(e._has_reset for e in self.envs):
            raise RuntimeError("Cannot call step() before calling reset()")
        acts, dts = [], []
        for e, a in zip(self.envs, actions):
            assert a.is_cuda, "mixed step takes device actions"
            if e._info.act_discrete and a.dtype not in (torch.int32, torch.int64, torch.uint8):
                a = a.to(torch.int32)
            if not e._info.act_discrete and a.dtype != torch.float32:
                a = a.to(torch.float32)
            a = a.contiguous()
            e._check_actions(e.num_envs, tuple(a.shape))
            acts.append(a)
            dts.append(_TORCH_ACT[a.dtype])
        ptrs = (ctypes.c_void_p * len(acts))(*[a.data_ptr() for a in acts])
        cdts = (ctypes.c_int * len(acts))(*dts)
        _native.check(self._lib.carlb_mixed_step(self._handles, ptrs, cdts, len(acts), self.envs[0]._stream()))
        outs = []
        for e in self.envs:
            state = e._add_context_to_state(e._obs)
            outs.append((state, e._reward, e._terminated_b, e._truncated_b,
                         {"context_id": e.context_id}))
        return outs
