"""Worker of tests/test_multigpu_gpu.py: one process per GPU (torchrun), sharded envs, the gathered
observation tensor (NCCL all-gather and the fused NVLink peer-store path) must equal what a single
GPU produces for the whole batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from carl_b200.envs import CARLBraxAnt, CARLCartPole, ContextTable
from carl_b200.parallel import ObsGather
from tests.util import sample_context_table
from oracle.classic import FEATURES


def say(rank, msg):
    print(f"[rank {rank}] {msg}", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 4096 + 3  # ragged over 2 ranks
    table = sample_context_table("cartpole", n, np.random.default_rng(0))
    ctxs = ContextTable(FEATURES["cartpole"], table)
    for mode in ("nccl", "fused"):
        say(rank, f"mode {mode}: build")
        ref = CARLCartPole(contexts=ctxs, device=dev, autoreset=True)
        env = CARLCartPole(contexts=ctxs, device=dev, autoreset=True, shard=(rank, world))
        g = ObsGather(env, mode=mode)
        o_ref, _ = ref.reset(seed=0)
        env.reset(seed=0)
        say(rank, f"mode {mode}: reset done, gathering")
        G = g.gather()
        assert torch.equal(G, o_ref["obs"]), f"{mode}: reset gather mismatch"
        say(rank, f"mode {mode}: reset gather ok")
        gen = torch.Generator(device="cpu").manual_seed(1)
        for t in range(25):
            a = torch.randint(0, 2, (n,), generator=gen, dtype=torch.int32).to(dev)
            o_ref, r_ref, *_ = ref.step(a)
            obs, r, te, tr, _ = env.step(a[env.env_lo:env.env_hi])
            G = g.gather()
            assert torch.equal(G, o_ref["obs"]), f"{mode}: step {t} gather mismatch"
            assert torch.equal(r, r_ref[env.env_lo:env.env_hi])
        say(rank, f"mode {mode}: steps ok")
        env.rollout(17, policy_seed=3)
        ref.rollout(17, policy_seed=3)
        G = g.gather()
        assert torch.equal(G, ref._obs), f"{mode}: rollout gather mismatch"
        # pipelined consumer: enqueue the next launch, then ask for the previous one (lag = 1)
        if mode == "fused" or (n % world == 0):
            prev = ref._obs.clone()
            env.rollout(5, policy_seed=4)
            ref.rollout(5, policy_seed=4)
            G1 = g.gather(lag=1)
            assert torch.equal(G1, prev), f"{mode}: lag-1 gather mismatch"
            assert torch.equal(g.gather(), ref._obs), f"{mode}: lag-0 after lag-1 mismatch"
        say(rank, f"mode {mode}: rollout ok")
        # masked reset: rows that are not reset must still reach the new slot
        mask_full = (np.arange(n) % 3 == 0)
        ref.reset(mask=mask_full)
        env.reset(mask=mask_full[env.env_lo:env.env_hi])
        G = g.gather()
        assert torch.equal(G, ref._obs), f"{mode}: masked reset gather mismatch"
        say(rank, f"mode {mode}: masked reset ok")
        torch.cuda.synchronize()
        dist.barrier()
        if mode == "fused":
            g.close()
    # Brax Ant, fused gather
    nb = 512
    refb = CARLBraxAnt(num_envs=nb, device=dev)
    envb = CARLBraxAnt(num_envs=nb, device=dev, shard=(rank, world))
    gb = ObsGather(envb, mode="fused")
    o_ref, _ = refb.reset(seed=5)
    envb.reset(seed=5)
    say(rank, "brax: reset done")
    assert torch.equal(gb.gather(), o_ref["obs"]), "brax reset gather mismatch"
    say(rank, "brax: reset gather ok")
    for t in range(5):
        a = (torch.rand(nb, 8, generator=torch.Generator().manual_seed(t)) * 2 - 1).to(dev)
        o_ref, *_ = refb.step(a)
        envb.step(a[envb.env_lo:envb.env_hi])
        assert torch.equal(gb.gather(), o_ref["obs"]), f"brax step {t} gather mismatch"
    torch.cuda.synchronize()
    dist.barrier()
    gb.close()
    dist.destroy_process_group()
    print(f"MGPU_OK rank {rank}")


if __name__ == "__main__":
    main()
