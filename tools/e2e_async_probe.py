#!/usr/bin/env python
"""Host-buffer (e2e) step of CARLCartPole, 65 536 contexts: synchronous env.step(numpy) with the completion-word
poll vs cudaStreamSynchronize, and the split-batch step_async / step_wait API with 1-4 parts. One JSON line."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def run_one(mode):
    import torch

    import bench
    from carl_b200 import hostmem
    from carl_b200.envs import CARLCartPole, ContextTable

    dev = torch.device("cuda", 0)
    n = bench.N_ENVS_PER_GPU
    names, table = bench.make_context_table(n)
    env = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True)
    env.reset(seed=0)
    acts = hostmem.pinned_empty((4, n), np.int32)
    acts[...] = np.random.default_rng(1).integers(0, 2, size=(4, n), dtype=np.int32)
    steps = 3000
    if mode.startswith("sync"):
        for w in range(50):
            env.step(acts[w % 4])
        acc = 0.0
        t0 = time.perf_counter()
        for j in range(steps):
            o, r, te, tr, _ = env.step(acts[j % 4])
            acc += float(r[0])
        dt = time.perf_counter() - t0
    else:
        k = int(mode.split("_")[1])
        env.async_parts = k
        bounds = [env.part_range(p) for p in range(k)]
        for w in range(20):
            env.step_async(acts[w % 4])
            env.step_wait()
        for p, (lo, hi) in enumerate(bounds):
            env.step_async(acts[0, lo:hi], part=p)
        acc = 0.0
        t0 = time.perf_counter()
        for j in range(steps):
            row = acts[(j + 1) % 4]
            for p, (lo, hi) in enumerate(bounds):
                st, r, te, tr, _ = env.step_wait(part=p)
                acc += float(r[0])  # the part's result is read on the host
                env.step_async(row[lo:hi], part=p)
        dt = time.perf_counter() - t0
        for p in range(k):
            env.step_wait(part=p)
    return {"us_per_step": dt / steps * 1e6, "env_steps_per_s": n * steps / dt}


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps(run_one(sys.argv[1])))
        sys.exit(0)
    out = {}
    for label, mode, env in [("sync_poll", "sync", {"CARLB_HOST_POLL": "1"}), ("sync_streamsync", "sync", {"CARLB_HOST_POLL": "0"}),
                             ("async_1", "async_1", {}), ("async_2", "async_2", {}), ("async_3", "async_3", {}),
                             ("async_4", "async_4", {}), ("async_8", "async_8", {})]:

        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), mode], capture_output=True, text=True, env=e, timeout=120)
        try:
            out[label] = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception:
            out[label] = {"error": (p.stderr or p.stdout)[-400:]}
    print(json.dumps(out))
