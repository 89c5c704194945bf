#!/usr/bin/env python
"""Short workloads for `ncu -k` captures (tools/gpu_visit.sh ncu): brax | brax_fma | step | step_host."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from carl_b200 import _native, hostmem


def main():
    what = sys.argv[1]
    dev = torch.device("cuda", 0)
    if what in ("brax", "brax_fma"):
        from carl_b200.envs import CARLBraxAnt

        n, T = 8192, 20
        ctxs = bench._brax_contexts(CARLBraxAnt, n, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)})
        env = CARLBraxAnt(contexts=ctxs, device=dev, context_mode="applied", arithmetic="fma" if what == "brax_fma" else "strict")
        env.reset(seed=0)
        tr = dict(obs=torch.empty(T, n, 27, device=dev), actions=torch.empty(T, n, 8, device=dev), reward=torch.empty(T, n, device=dev),
                  done=torch.empty(T, n, dtype=torch.uint8, device=dev))
        traj = _native.Traj(obs=tr["obs"].data_ptr(), actions=tr["actions"].data_ptr(), reward=tr["reward"].data_ptr(), done=tr["done"].data_ptr())
        for j in range(8):
            _native.check(env._lib.carlb_env_rollout(env._handle, T, 7, j * T, None, _native.ACT_F32, ctypes.byref(traj),
                                                     torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize()
        return
    from carl_b200.envs import CARLCartPole, ContextTable

    n = bench.N_ENVS_PER_GPU
    names, table = bench.make_context_table(n)
    env = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True)
    env.reset(seed=0)
    if what == "step":
        a = torch.randint(0, 2, (8, n), dtype=torch.int32, device=dev)
        for j in range(16):
            env.step(a[j % 8])
    else:
        acts = hostmem.pinned_empty((4, n), np.int32)
        acts[...] = np.random.default_rng(1).integers(0, 2, size=(4, n), dtype=np.int32)
        for j in range(16):
            env.step(acts[j % 4])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
