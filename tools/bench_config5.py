"""BASELINE config 5: CARLHalfcheetah + CARLHopper, 16 384 contexts (8 192 each), sharded over the GPUs
of one box with the per-step observation all-gather. One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_config5.py

STRONG scaling (the 16 384 contexts are split over the ranks). Every env-step launches both bodies' step
kernels on the rank's shard and gathers both observation tensors on every rank (fused NVLink stores or
NCCL); device time is the max over ranks. Prints one JSON line per gather mode."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from carl_b200.context import ContextSampler, UniformFloatContextFeature
from carl_b200.envs import CARLBraxHalfcheetah, CARLBraxHopper, ContextTable
from carl_b200.parallel import ObsGather


def table(cls, feats, n):
    names = list(cls.get_context_space().get_default_context().keys())
    s = ContextSampler([UniformFloatContextFeature(k, lo, hi) for k, (lo, hi) in feats.items()],
                       context_space=cls.get_context_space(), seed=0)
    return ContextTable(names, s.sample_context_table(n, names))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29555",
                            rank=rank, world_size=world)
    n, steps, warm = 8192, int(os.environ.get("C5_STEPS", 300)), 30
    feats = {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}
    for mode in ("fused", "nccl"):
        envs = [cls(contexts=table(cls, feats, n), device=dev, context_mode="applied", shard=(rank, world))
                for cls in (CARLBraxHalfcheetah, CARLBraxHopper)]
        gathers = [ObsGather(e, mode=mode) for e in envs]
        acts = []
        for e in envs:
            e.reset(seed=0)
            gen = torch.Generator(device="cpu").manual_seed(1 + rank)
            acts.append((torch.rand(e.num_envs, e._info.act_dim, generator=gen) * 2 - 1).to(dev))
        for g in gathers:
            g.gather()
        torch.cuda.synchronize(dev)

        def one_step():
            for e, a in zip(envs, acts):
                e.step(a)
            return [g.gather() for g in gathers]

        for _ in range(warm):
            one_step()
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            G = one_step()
        e1.record()
        torch.cuda.synchronize(dev)
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
        ok = all(torch.isfinite(x).all().item() for x in G) and G[0].shape == (n, 17) and G[1].shape == (n, 11)
        if rank == 0:
            print(json.dumps({"workload": "CARLBraxHalfcheetah 8192 + CARLBraxHopper 8192 contexts, sharded, obs all-gather every step",
                              "n_gpus": world, "gather": mode, "scaling": "strong", "env_steps_per_s": 2 * n / (ms * 1e-3),
                              "us_per_step": ms * 1e3, "steps": steps, "gathered_shapes_ok": bool(ok)}), flush=True)
        for g in gathers:
            g.close()
        del envs, gathers
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
