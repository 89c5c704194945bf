"""Per-step and fused Brax throughput vs batch size for the current CARLB_BRAX_PACK setting (run once per
setting: the library reads the variable once). Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import carl_b200.envs as E


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"pack": os.environ.get("CARLB_BRAX_PACK", "default")}
for name in ("CARLBraxAnt", "CARLBraxHalfcheetah", "CARLBraxHopper"):
    for n in (512, 1024, 2048, 4096, 8192):
        env = getattr(E, name)(num_envs=n)
        env.reset(seed=0)
        a = torch.rand(n, env._info.act_dim, device="cuda") * 2 - 1
        ms1 = timed(lambda: env.step(a), 100)
        T = 20
        msT = timed(lambda: env.rollout(T, policy_seed=1, record=False), 10, warm=2)
        out[f"{name}_{n}"] = {"step_us": round(ms1 * 1e3, 1), "fused_us_per_step": round(msT * 1e3 / T, 1)}
print(json.dumps(out))
