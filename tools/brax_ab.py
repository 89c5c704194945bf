"""A/B helper for Brax kernel work: fused-rollout throughput of the locomotion bodies (8 192 contexts, 20 steps per
launch, strict and FMA arithmetic) with whatever libcarlb.so is in place. `--ncu BODY ARITH` runs a few launches of one
body only (for `ncu --metrics smsp__inst_executed.sum`)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import carl_b200.envs as E
from tools.bench_extras import table, timed

BODIES = {"ant": "CARLBraxAnt", "halfcheetah": "CARLBraxHalfcheetah", "hopper": "CARLBraxHopper", "walker2d": "CARLBraxWalker2d",
          "humanoid": "CARLBraxHumanoid"}
FEATS = {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}


def make(body, arithmetic, n=8192):
    cls = getattr(E, BODIES[body])
    env = cls(contexts=table(cls, FEATS, n), context_mode="applied", arithmetic=arithmetic)
    env.reset(seed=0)
    return env


def main():
    if "--ncu" in sys.argv:
        body, arith = sys.argv[sys.argv.index("--ncu") + 1:sys.argv.index("--ncu") + 3]
        env = make(body, arith)
        for _ in range(3):
            env.rollout(20, policy_seed=1, record=False)
        torch.cuda.synchronize()
        return
    out = {}
    n, T = 8192, 20
    for body in BODIES:
        for arith in ("strict", "fma"):
            env = make(body, arith, n)
            ms = min(timed(lambda: env.rollout(T, policy_seed=1, record=False), 10, warm=2) for _ in range(3))
            out[f"{body}_{arith}"] = round(n * T / (ms * 1e-3) / 1e6, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
