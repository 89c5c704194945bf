#!/usr/bin/env python
"""ONE GPU, world size 1: the fused-rollout train of bench.py with and without a fused gather attached (the gather
then pushes to its own buffer), to separate the cost of the in-kernel gather machinery (publisher warp, barriers,
counters) from NVLink effects. Usage: gather_probe_1gpu.py [none|pipelined|sync] [--ncu]"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from carl_b200 import _native
from carl_b200.envs import CARLCartPole, ContextTable
from carl_b200.parallel import ObsGather


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "none"
    short = "--ncu" in sys.argv
    dev = torch.device("cuda", 0)
    n, T = bench.N_ENVS_PER_GPU, int(os.environ.get("PROBE_T", "20"))
    names, table = bench.make_context_table(n)
    env = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True)
    env.reset(seed=0)
    g = None
    if mode != "none":
        g = ObsGather(env, mode="fused", pipelined=(mode == "pipelined"), symmetric="ipc")
    n_slots = max(2, int(np.ceil(1.5 * bench.L2_BYTES / (T * n * bench.TRAJ_BYTES))) + 1)
    ring = [dict(obs=torch.empty(T, n, 4, device=dev), actions=torch.empty(T, n, dtype=torch.int32, device=dev),
                 reward=torch.empty(T, n, device=dev), done=torch.empty(T, n, dtype=torch.uint8, device=dev)) for _ in range(n_slots)]
    trajs = [_native.Traj(obs=r["obs"].data_ptr(), actions=r["actions"].data_ptr(), reward=r["reward"].data_ptr(),
                          done=r["done"].data_ptr()) for r in ring]
    stream = torch.cuda.current_stream(dev)
    cnt = [0]

    def launch(st):
        j = cnt[0]
        cnt[0] += 1
        _native.check(env._lib.carlb_env_rollout(env._handle, T, 12345, j * T, None, _native.ACT_I32, ctypes.byref(trajs[j % n_slots]), st))

    for _ in range(8):
        launch(stream.cuda_stream)
    torch.cuda.synchronize()
    if short:
        for _ in range(8):
            launch(stream.cuda_stream)
        torch.cuda.synchronize()
        return
    side = torch.cuda.Stream(dev)
    side.wait_stream(stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(32):
            launch(torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(300):
        graph.replay()
    e1.record(stream)
    torch.cuda.synchronize()
    check = None
    if g is not None:  # the gathered tensor must be the handle's observation (lag 1 in pipelined mode, lag 0 in sync mode)
        g.resync()
        prev = env._obs.clone()
        launch(stream.cuda_stream)
        torch.cuda.synchronize()
        check = bool(torch.equal(g.gather(lag=1), prev)) and bool(torch.equal(g.gather(lag=0), env._obs))
        assert check, "gathered tensor != observation"
    print(json.dumps({"mode": mode, "gather_ok": check, "debug": os.environ.get("CARLB_GATHER_DEBUG"), "block": os.environ.get("CARLB_ROLLOUT_BLOCK"), "pdl": os.environ.get("CARLB_ROLLOUT_PDL"), "T": T,
                      "us_per_launch": e0.elapsed_time(e1) * 1e3 / (300 * 32)}))


if __name__ == "__main__":
    main()
