#!/usr/bin/env python
"""Per-launch cost of the fused CartPole rollout kernel as a function of the fused step count T.

For every T: (a) R back-to-back launches on one stream, bracketed by two events (eager enqueue);
(b) the same launches replayed from a CUDA graph; (c) ONE launch on an idle stream between two events
(what a `--steps T` bench with a single launch in its timed region sees). Separates the per-launch
fixed cost (launch latency, prologue, tail) from the per-step cost. Output: one JSON line.
"""
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


def main():
    dev = torch.device("cuda", 0)
    n = bench.N_ENVS_PER_GPU
    names, table = bench.make_context_table(n)
    env = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True)
    env.reset(seed=0)
    lib, handle = env._lib, env._handle
    stream = torch.cuda.current_stream(dev)
    out = {}
    for T in [1, 2, 5, 10, 20, 50, 100, 500]:
        slot_bytes = T * n * bench.TRAJ_BYTES
        n_slots = max(2, int(np.ceil(1.5 * bench.L2_BYTES / slot_bytes)) + 1)
        n_slots = min(n_slots, 256)
        ring = [dict(obs=torch.empty(T, n, 4, device=dev), actions=torch.empty(T, n, dtype=torch.int32, device=dev),
                     reward=torch.empty(T, n, device=dev), done=torch.empty(T, n, dtype=torch.uint8, device=dev))
                for _ in range(n_slots)]
        trajs = [_native.Traj(obs=r["obs"].data_ptr(), actions=r["actions"].data_ptr(), reward=r["reward"].data_ptr(),
                              done=r["done"].data_ptr()) for r in ring]

        def launch(j, st):
            _native.check(lib.carlb_env_rollout(handle, T, 12345, j * T, None, _native.ACT_I32,
                                                ctypes.byref(trajs[j % n_slots]), st))

        for j in range(n_slots):
            launch(j, stream.cuda_stream)
        torch.cuda.synchronize()
        R = int(max(50, min(20000, 0.05 / (2e-6 + T * 0.45e-6))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for j in range(R):
            launch(j, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        eager_us = e0.elapsed_time(e1) * 1e3 / R
        # graph of G launches
        G = n_slots * max(1, 32 // n_slots)
        side = torch.cuda.Stream(dev)
        side.wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for j in range(G):
                launch(j, torch.cuda.current_stream(dev).cuda_stream)
        graph.replay()
        torch.cuda.synchronize()
        reps = max(3, R // G)
        e0.record(stream)
        for _ in range(reps):
            graph.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) * 1e3 / (reps * G)
        single = []
        for j in range(12):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            launch(j, stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            single.append(a.elapsed_time(b) * 1e3)
        out[T] = {"eager_us_per_launch": eager_us, "graph_us_per_launch": graph_us,
                  "single_idle_stream_us_median": float(np.median(single[2:])), "launches": R, "ring_slots": n_slots,
                  "env_steps_per_s_graph": n * T / (graph_us * 1e-6)}
        del ring, trajs, graph
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
