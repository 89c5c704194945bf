"""Worker of tests/test_multigpu_gpu.py: one process per GPU (torchrun), sharded envs, the gathered
observation tensor (NCCL all-gather and the fused NVLink push path in all its variants) must equal what a
single GPU produces for the whole batch -- step by step, with ragged shards, masked resets, rank skew,
pipelined (lag-1) consumers and CUDA-graph replays."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from carl_b200.envs import CARLBraxAnt, CARLCartPole, CARLPendulum, ContextTable
from carl_b200.parallel import ObsGather
from tests.util import sample_context_table
from oracle.classic import FEATURES


def say(rank, msg):
    print(f"[rank {rank}] {msg}", flush=True)


def skew(rank, t):
    """Delay one rank's stream by ~0.2-0.6 ms at irregular launches (a different rank each time)."""
    if (t * 7 + 3) % 5 == 0 and (t // 5) % dist.get_world_size() == rank:
        torch.cuda._sleep(400_000 + 100_000 * (t % 9))


def classic_variant(rank, world, dev, n, label, mode, pipelined, symmetric):
    table = sample_context_table("cartpole", n, np.random.default_rng(0))
    ctxs = ContextTable(FEATURES["cartpole"], table)
    say(rank, f"{label}: build")
    ref = CARLCartPole(contexts=ctxs, device=dev, autoreset=True)
    env = CARLCartPole(contexts=ctxs, device=dev, autoreset=True, shard=(rank, world))
    g = ObsGather(env, mode=mode, pipelined=pipelined, symmetric=symmetric) if mode == "fused" else ObsGather(env, mode=mode)
    say(rank, f"{label}: transport {g.transport}")
    o_ref, _ = ref.reset(seed=0)
    env.reset(seed=0)
    assert torch.equal(g.gather(), o_ref["obs"]), f"{label}: reset gather mismatch"
    gen = torch.Generator(device="cpu").manual_seed(1)
    prev = o_ref["obs"].clone()
    for t in range(25):
        a = torch.randint(0, 2, (n,), generator=gen, dtype=torch.int32).to(dev)
        o_ref, r_ref, *_ = ref.step(a)
        skew(rank, t)
        obs, r, te, tr, _ = env.step(a[env.env_lo:env.env_hi])
        if t % 2 == 0 or mode == "nccl":
            assert torch.equal(g.gather(), o_ref["obs"]), f"{label}: step {t} gather mismatch"
        elif mode == "fused":
            assert torch.equal(g.gather(lag=1), prev), f"{label}: step {t} lag-1 gather mismatch"
        assert torch.equal(r, r_ref[env.env_lo:env.env_hi])
        prev = o_ref["obs"].clone()
    say(rank, f"{label}: steps ok")
    env.rollout(17, policy_seed=3)
    ref.rollout(17, policy_seed=3)
    assert torch.equal(g.gather(), ref._obs), f"{label}: rollout gather mismatch"
    if mode == "fused" or (n % world == 0):
        # pipelined consumer over many launches with rank skew: enqueue launch k+1, then take obs k (lag = 1)
        prev = ref._obs.clone()
        for t in range(60):
            skew(rank, t)
            env.rollout(3, policy_seed=4, step_base=3 * t)
            ref.rollout(3, policy_seed=4, step_base=3 * t)
            G1 = g.gather(lag=1)
            assert torch.equal(G1, prev), f"{label}: lag-1 gather mismatch at launch {t}"
            prev = ref._obs.clone()
        assert torch.equal(g.gather(), ref._obs), f"{label}: lag-0 after lag-1 mismatch"
    say(rank, f"{label}: rollout + skewed lag-1 ok")
    mask_full = (np.arange(n) % 3 == 0)
    ref.reset(mask=mask_full)
    env.reset(mask=mask_full[env.env_lo:env.env_hi])
    assert torch.equal(g.gather(), ref._obs), f"{label}: masked reset gather mismatch"
    say(rank, f"{label}: masked reset ok")
    if mode == "fused":
        # CUDA-graph replay of obs-producing launches (slot / flag value come from device memory)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        L, reps = 6, 5
        with torch.cuda.graph(graph, stream=side):
            for j in range(L):
                env.rollout(2, policy_seed=9, step_base=2 * j)
        torch.cuda.synchronize(dev)
        for _ in range(reps):
            graph.replay()
        g.resync()
        for _ in range(reps):
            for j in range(L):
                ref.rollout(2, policy_seed=9, step_base=2 * j)
        torch.cuda.synchronize(dev)
        assert torch.equal(env._obs, ref._obs[env.env_lo:env.env_hi]), f"{label}: graph replay state mismatch"
        assert torch.equal(g.gather(), ref._obs), f"{label}: gather after graph replay mismatch"
        say(rank, f"{label}: graph replay ok")
    torch.cuda.synchronize()
    dist.barrier()
    if mode == "fused":
        g.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    only = os.environ.get("CARLB_MGPU_ONLY", "")  # "brax": just the Brax part (a quick check of the Brax kernels' pushes)
    if only != "brax":
        classic_part(rank, world, dev)
    brax_part(rank, world, dev)
    dist.destroy_process_group()
    print(f"MGPU_OK rank {rank}")


def classic_part(rank, world, dev):
    n = 4096 + 3  # ragged over the ranks
    classic_variant(rank, world, dev, n, "nccl", "nccl", False, "ipc")
    for symmetric in ("ipc", "auto"):
        for pipelined in (False, True):
            classic_variant(rank, world, dev, n, f"fused/{symmetric}/{'pipelined' if pipelined else 'sync'}", "fused",
                            pipelined, symmetric)
    # BASELINE-size shards (1 024 CTAs per launch: every CTA arrives with a device-scope release, ONE system fence
    # per launch must cover all of their peer stores)
    big = 65536 + 5
    classic_variant(rank, world, dev, big, "big/fused/auto/pipelined", "fused", True, "auto")
    classic_variant(rank, world, dev, big, "big/fused/ipc/sync", "fused", False, "ipc")
    # an odd obs width (Pendulum: 3 floats per row -> scalar / 8-byte row stores)
    npd = 1024 + 1
    tp = sample_context_table("pendulum", npd, np.random.default_rng(1))
    cp = ContextTable(FEATURES["pendulum"], tp)
    for pipelined in (False, True):
        refp = CARLPendulum(contexts=cp, device=dev, autoreset=True)
        envp = CARLPendulum(contexts=cp, device=dev, autoreset=True, shard=(rank, world))
        gp = ObsGather(envp, mode="fused", pipelined=pipelined)
        refp.reset(seed=2)
        envp.reset(seed=2)
        for t in range(6):
            a = (torch.rand(npd, 1, generator=torch.Generator().manual_seed(t)) * 4 - 2).to(dev)
            refp.step(a)
            envp.step(a[envp.env_lo:envp.env_hi])
            assert torch.equal(gp.gather(), refp._obs), f"pendulum step {t} gather mismatch"
        torch.cuda.synchronize()
        dist.barrier()
        gp.close()
    say(rank, "pendulum ok")


def brax_part(rank, world, dev):
    # Brax Ant, fused gather (immediate pushes; pipelined = one push behind)
    nb = 512
    for pipelined in (False, True):
        refb = CARLBraxAnt(num_envs=nb, device=dev)
        envb = CARLBraxAnt(num_envs=nb, device=dev, shard=(rank, world))
        gb = ObsGather(envb, mode="fused", pipelined=pipelined)
        o_ref, _ = refb.reset(seed=5)
        envb.reset(seed=5)
        assert torch.equal(gb.gather(), o_ref["obs"]), "brax reset gather mismatch"
        prev = o_ref["obs"].clone()
        for t in range(6):
            a = (torch.rand(nb, 8, generator=torch.Generator().manual_seed(t)) * 2 - 1).to(dev)
            o_ref, *_ = refb.step(a)
            skew(rank, t)
            envb.step(a[envb.env_lo:envb.env_hi])
            if t % 2 == 0:
                assert torch.equal(gb.gather(), o_ref["obs"]), f"brax step {t} gather mismatch"
            else:
                assert torch.equal(gb.gather(lag=1), prev), f"brax step {t} lag-1 gather mismatch"
            prev = o_ref["obs"].clone()
        torch.cuda.synchronize()
        dist.barrier()
        gb.close()
    say(rank, "brax ok")


if __name__ == "__main__":
    main()
