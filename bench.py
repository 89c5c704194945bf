#!/usr/bin/env python
"""Headline benchmark: env-steps/sec of the batched-step hot path.

Workload (BASELINE.json configs[1]): CARLCartPole, 65 536 contexts per GPU sampled
by ``ContextSampler(seed=0)`` with gravity~U(5,15), length~U(0.25,1.0), masscart~U(0.5,2.0);
synthetic uniform random-policy actions; auto-reset on; TimeLimit 500. A bench "step" is ONE
env-step of every env instance in the batch.

Three measurements go into the one JSON line:

* ``value``  -- device-resident whole-job throughput. The K timed steps run as K/T launches of the
  fused T-step rollout kernel (``carlb_env_rollout``: state in registers, in-kernel Philox policy),
  each launch streaming its full trajectory (obs, action, reward, done = 25 B per env-step) into a
  ring of HBM buffers larger than L2, so every step's outputs really go to DRAM.
* ``step_api`` -- the same workload through one ``carlb_env_step`` launch per step (the
  reference's ``CARLEnv.step`` contract, 90 algorithmic bytes per env-step), device-resident
  actions streamed from a ring larger than L2, launches replayed from a CUDA graph.
* ``e2e`` -- the gym-style call a user makes, ``env.step(numpy_actions)``: HOST buffers in and out,
  host->device and device->host copies inside the timed region (``carlb_env_step_host``).

``--impl reference`` times the CPU oracle port of the reference's step (the reference itself is
pure Python over gymnasium, which is not installable here) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ENVS_PER_GPU = 65536
METRIC = "env-steps/sec at N contexts"
WORKLOAD = (f"CARLCartPole, {N_ENVS_PER_GPU} sampled contexts (gravity/length/masscart) per GPU (BASELINE configs[1]), "
            f"uniform random policy, autoreset, TimeLimit 500")
UNIT = "env-steps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
L2_BYTES = 126 * 1024 * 1024

# algorithmic bytes per env-step
STEP_CONTRACT_BYTES = 90  # R: state16+ctx24+action4+elapsed4 ; W: state16+obs16+reward4+flags2+elapsed4
TRAJ_BYTES = 16 + 4 + 4 + 1  # fused rollout: obs + action + reward + done written per env-step


def make_context_table(n: int):
    from carl_b200.context import ContextSampler, UniformFloatContextFeature
    from carl_b200.envs import CARLCartPole

    names = list(CARLCartPole.get_context_space().get_default_context().keys())
    sampler = ContextSampler(
        [UniformFloatContextFeature("gravity", 5, 15), UniformFloatContextFeature("length", 0.25, 1.0),
         UniformFloatContextFeature("masscart", 0.5, 2.0)],
        context_space=CARLCartPole.get_context_space(), seed=0)
    return names, sampler.sample_context_table(n, names)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_profile(rep: str, fused_steps: int | None = None):
    """Counters of the dominant kernel from the newest committed ncu --set full summary
    (profiles/<tag>_traffic.json, written by tools/summarize_ncu.py) -- of a capture that fused `fused_steps`
    env-steps per launch when given: (dict, file name) or ({}, None)."""
    d = os.path.join(ROOT, "profiles")
    try:
        files = sorted((f for f in os.listdir(d) if f.endswith("_traffic.json")), reverse=True)
        for f in files:
            j = json.load(open(os.path.join(d, f)))
            if rep in j and (fused_steps is None or j[rep].get("fused_steps_per_launch") == fused_steps):
                return dict(j[rep]), f
    except Exception:
        pass
    return {}, None


def issue_roof(rep: str, env_steps_per_launch: int, env_steps_per_s: float, sm_mhz: float | None, n_sms: int = 148):
    """Instruction-issue roofline of an issue-bound kernel (the Brax steps): warp instructions per env-step from the
    newest committed ncu capture of `rep` (profiles/<tag>_traffic.json: smsp__inst_executed.sum per launch) against the
    chip's issue rate, 4 schedulers per SM x one warp instruction per cycle at the SM clock sampled during the run.
    None when no capture is committed. Never raises (the bench line must not depend on it)."""
    try:
        prof, f = measured_profile(rep)
        wi = float(prof["warp_instructions_per_launch"]) / float(env_steps_per_launch)
        mhz = float(sm_mhz) if sm_mhz else 1965.0
        peak = n_sms * 4 * mhz * 1e6 / wi  # env-steps/s at one issued warp instruction per scheduler and cycle
        return {"bound": "instruction issue", "warp_instructions_per_env_step": wi, "peak_env_steps_per_s": peak,
                "frac": env_steps_per_s / peak, "sm_mhz": mhz,
                "source": f"profiles/{f} [{rep}]: smsp__inst_executed.sum of one launch of {env_steps_per_launch} env-steps; "
                          f"peak = {n_sms} SMs x 4 schedulers x SM clock / instructions per env-step"}
    except Exception:
        return None


def measured_traffic(rep: str):
    p, f = measured_profile(rep)
    return (float(p["dram_bytes_per_launch"]) if "dram_bytes_per_launch" in p else None), f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_baseline_leg(steps_per_call: int, target_seconds: float, n_envs: int = N_ENVS_PER_GPU):
    """The oracle port of the reference's CartPole step, all host threads, bounded sample."""
    import ctypes

    import oracle

    L = oracle.lib()
    L.oracle_cartpole_rollout_baseline.restype = ctypes.c_longlong
    L.oracle_cartpole_rollout_baseline.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
        ctypes.c_int, ctypes.c_void_p]
    threads = int(L.oracle_max_threads())
    quota = cgroup_cpu_quota()
    if quota:  # more runnable threads than granted CPUs only get throttled (see run_reference_arm)
        threads = max(1, min(threads, int(round(quota))))
    _, table = make_context_table(n_envs)
    table = np.ascontiguousarray(table)
    state = np.random.default_rng(0).uniform(-0.1, 0.1, (n_envs, 4))

    def run(steps):
        ret = ctypes.c_double()
        t0 = time.perf_counter()
        done = L.oracle_cartpole_rollout_baseline(n_envs, steps, state.ctypes.data, table.ctypes.data, 500, 0, 0,
                                                  threads, ctypes.byref(ret))
        return done, time.perf_counter() - t0

    run(2)  # warm
    run(steps_per_call)
    total, elapsed, reps = 0, 0.0, 0
    while elapsed < target_seconds:
        d, t = run(steps_per_call)
        total += d
        elapsed += t
        reps += 1
    return {"value": total / elapsed, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_envs} CARLCartPole contexts x {steps_per_call * reps} steps "
                      f"({total:.3g} env-steps, {elapsed:.1f} s) -- C/OpenMP oracle port of the reference step",
            "threads": threads}


def brax_cpu_baseline(sysd, ctx, target_seconds: float = 3.0, threads: int | None = None):
    """The Brax oracle port (plain C, OpenMP over the env instances) stepping the same batch with a uniform random
    policy on the host cores: the CPU number next to `ant_8192` (bounded sample)."""
    import ctypes

    import oracle
    from carl_b200.envs import brax_system as bs
    from oracle.brax import OracleBraxEnv

    L = oracle.lib()
    if threads is None:
        threads = int(L.oracle_max_threads())
        quota = cgroup_cpu_quota()
        if quota:
            threads = max(1, min(threads, int(round(quota))))
    # the OpenMP thread count is process-wide: set it through the classic baseline entry point (0 steps)
    L.oracle_cartpole_rollout_baseline.restype = ctypes.c_longlong
    L.oracle_cartpole_rollout_baseline.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
        ctypes.c_int, ctypes.c_void_p]
    ret = ctypes.c_double()
    dummy_state, dummy_table = np.zeros((1, 4)), np.ascontiguousarray(make_context_table(1)[1])
    L.oracle_cartpole_rollout_baseline(1, 0, dummy_state.ctypes.data, dummy_table.ctypes.data, 500, 0, 0, threads, ctypes.byref(ret))
    n = ctx.shape[0]
    rng = np.random.default_rng(3)
    env = OracleBraxEnv(sysd, ctx, autoreset=True)
    init_q = sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + sysd["n_q"]]
    q = (init_q[None] + rng.uniform(-0.1, 0.1, (n, sysd["n_q"]))).astype(np.float32)
    qd = (0.1 * rng.standard_normal((n, sysd["n_qd"]))).astype(np.float32)
    env.init_from_q(q, qd)
    acts = rng.uniform(-1, 1, (8, n, sysd["n_act"])).astype(np.float32)
    env.step(acts[0])  # warm
    steps, elapsed = 0, 0.0
    while elapsed < target_seconds:
        t0 = time.perf_counter()
        env.step(acts[steps % 8])
        elapsed += time.perf_counter() - t0
        steps += 1
    return {"value": n * steps / elapsed, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} envs x {steps} env-steps ({elapsed:.1f} s) -- C/OpenMP oracle port of the Brax spring step"}


def cgroup_cpu_quota():
    """CPUs granted by the container's CFS quota (cgroup v2 `cpu.max`), or None."""
    try:
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        return None if q == "max" else float(q) / float(period)
    except Exception:
        pass
    try:  # cgroup v1
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        period = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        return None if q <= 0 else q / period
    except Exception:
        return None


def try_real_reference(seconds: float = 2.0):
    """B0: the REAL reference -- `carl.envs.CARLCartPole().step` over gymnasium -- if it can be
    imported on this box (it cannot in this image: profiles/r02a_pip_install_attempt.txt). Returns a dict or the
    import error."""
    try:
        import gymnasium  # noqa: F401
        sys.path.insert(0, "/root/reference")
        from carl.envs import CARLCartPole as RefCartPole
    except Exception as e:  # the expected outcome here
        return {"available": False, "why": f"{type(e).__name__}: {e}"}
    env = RefCartPole()
    env.reset(seed=0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, _, te, tr, _ = env.step(env.action_space.sample())
        if te or tr:
            env.reset()
        n += 1
    return {"available": True, "value": n / (time.perf_counter() - t0), "unit": UNIT, "cores": 1}


def _b2_worker(n_steps):
    from oracle.classic import scalar_python_cartpole_steps_per_s

    return scalar_python_cartpole_steps_per_s(n_steps)


def python_scalar_all_cores(n_steps: int = 60_000):
    """B2: the scalar-Python restatement of the reference's per-step cost shape (one env object,
    math.sin/cos, TimeLimit, a fresh {"obs","context"} dict per step) in one process per host core."""
    import multiprocessing as mp

    procs = max(1, len(os.sched_getaffinity(0)))
    quota = cgroup_cpu_quota()
    if quota:
        procs = max(1, min(procs, int(round(quota))))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        rates = pool.map(_b2_worker, [n_steps] * procs)
    wall = time.perf_counter() - t0
    return {"value": float(sum(rates)), "unit": UNIT, "cores": procs, "per_core": float(np.mean(rates)),
            "note": f"sum of the per-process rates of {procs} processes x {n_steps} steps (pool wall time {wall:.1f} s incl. spawn)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import ctypes

    import oracle

    L = oracle.lib()
    L.oracle_cartpole_rollout_baseline.restype = ctypes.c_longlong
    L.oracle_cartpole_rollout_baseline.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
        ctypes.c_int, ctypes.c_void_p]
    # torchrun exports OMP_NUM_THREADS=1 to its workers when N > 1; the CPU arm must not inherit that
    threads = max(int(L.oracle_max_threads()), len(os.sched_getaffinity(0)))
    n = N_ENVS_PER_GPU * args.gpus
    _, table = make_context_table(n)
    table = np.ascontiguousarray(table)
    state = np.random.default_rng(0).uniform(-0.1, 0.1, (n, 4))
    ret = ctypes.c_double()
    # The CPU arm is given every advantage: each thread owns a slice of the envs and runs all K steps
    # without a barrier; the thread count (all logical CPUs, or half of them = one per physical core
    # on SMT hosts) is whichever calibrates faster; the pool is warm; and because K steps of 65 536
    # envs are only a few milliseconds of CPU work, the K-step pass is repeated until ~2 s are timed.
    def timed(k, nthreads):
        t0 = time.perf_counter()
        d = L.oracle_cartpole_rollout_baseline(n, k, state.ctypes.data, table.ctypes.data, 500, 0, 0, nthreads,
                                               ctypes.byref(ret))
        return d, time.perf_counter() - t0

    cal = {}
    candidates = {threads, max(1, threads // 2)}
    quota = cgroup_cpu_quota()
    if quota:  # a CFS quota throttles sustained runs: one thread per granted CPU is a further candidate
        candidates.add(max(1, min(threads, int(round(quota)))))
    for nt in sorted(candidates):
        timed(50, nt)
        d_sum, t_sum = 0, 0.0
        while t_sum < 0.6:  # long enough for a CFS quota (100 ms periods) to show
            d, t = timed(300, nt)
            d_sum += d
            t_sum += t
        cal[nt] = d_sum / t_sum
    threads = max(cal, key=cal.get)
    for _ in range(max(3, args.warmup // max(1, args.steps))):
        timed(args.steps, threads)
    done, dt, reps = 0, 0.0, 0
    while dt < 2.0 or reps < 3:
        d, t = timed(args.steps, threads)
        done += d
        dt += t
        reps += 1
    v = done / dt
    dt = dt / reps  # seconds per K-step pass
    from oracle.classic import scalar_python_cartpole_steps_per_s

    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_envs": n, "policy": "uniform random", "autoreset": True, "time_limit": 500,
                   "passes": reps, "policy_rng": "xorshift (CPU arm)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} contexts x {args.steps} steps x {reps} passes, C/OpenMP oracle port "
                                   f"(gymnasium/CARL not installable); thread calibration {cal}; cgroup cpu quota {quota}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "python_scalar_reference_shape": {
            "value": scalar_python_cartpole_steps_per_s(100_000), "unit": UNIT, "cores": 1,
            "note": "B1: one env, scalar float64 Python + TimeLimit + dict obs per step: the reference's actual cost shape"},
        "real_reference_B0": try_real_reference(),
        "gpu_launches": 0,
    }
    try:
        out["python_scalar_all_cores_B2"] = python_scalar_all_cores()
    except Exception as e:  # pragma: no cover - never allowed to break the line
        out["python_scalar_all_cores_B2"] = {"error": repr(e)}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------ GPU arm
def timed_train(replay, n_per_replay: int, stream, barrier, min_gpu_s: float, min_replays: int = 5, max_replays: int = 4000):
    """Times a train of back-to-back `replay()` calls (each enqueues `n_per_replay` units of work on `stream`)
    long enough for >= min_gpu_s of GPU time: one event before, one after every replay. Returns
    (train_ms, [ms per replay], replays). The estimate for the train length comes from one untimed replay."""
    import torch

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    replay()
    e1.record(stream)
    barrier()
    est_ms = max(e0.elapsed_time(e1), 1e-3)
    reps = int(min(max_replays, max(min_replays, math.ceil(min_gpu_s * 1e3 / est_ms))))
    # every rank must enqueue the same number of launches (the gather is a collective)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([reps], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reps = int(t.item())
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    barrier()
    evs[0].record(stream)
    for r in range(reps):
        replay()
        evs[r + 1].record(stream)
    barrier()
    per = [evs[r].elapsed_time(evs[r + 1]) for r in range(reps)]
    return evs[0].elapsed_time(evs[reps]), per, reps


def max_over_ranks(x: float, dev, distributed: bool) -> float:
    import torch
    import torch.distributed as dist

    if not distributed:
        return float(x)
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def tensor_sha(t) -> str:
    import hashlib

    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def run_gpu_arm(args):
    import ctypes

    import torch
    import torch.distributed as dist

    from carl_b200 import _native, hostmem
    from carl_b200.envs import CARLCartPole, ContextTable

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        # stdout carries exactly one JSON line: NCCL's version banner / debug output goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/carlb_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    n_local = N_ENVS_PER_GPU
    n_global = n_local * world
    names, table = make_context_table(n_global)

    def make_env():
        e = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True, shard=(rank, world))
        e.reset(seed=0)
        return e

    env = make_env()
    assert env.num_envs == n_local
    gather = None
    gather_mode = args.gather
    if distributed:
        from carl_b200.parallel import ObsGather

        if gather_mode == "fused":
            # every rank must succeed in mapping its peers; otherwise all fall back to NCCL
            ok = 1
            try:
                gather = ObsGather(env, mode="fused", pipelined=True)
            except Exception as e:  # pragma: no cover - depends on the box
                ok = 0
                print(f"[rank {rank}] fused gather unavailable ({e}); falling back to NCCL", file=sys.stderr)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if gather is not None:
                    gather.close()
                gather, gather_mode = None, "nccl"
                env = make_env()  # a handle that had a gather attached is rebuilt without one
        if gather is None:
            gather = ObsGather(env, mode="nccl")
    fused_gather = gather is not None and gather.mode == "fused"

    K, W = args.steps, args.warmup
    T = max(1, min(K, args.fuse))  # fused steps per launch
    plan = [T] * (K // T) + ([K % T] if K % T else [])  # the launches of ONE pass = exactly K steps
    info = env._info
    # trajectory ring: outputs larger than L2 so every launch's writes go to DRAM
    slot_bytes = T * n_local * TRAJ_BYTES
    n_slots = max(2, int(math.ceil(1.5 * L2_BYTES / slot_bytes)) + 1)
    ring = [dict(obs=torch.empty(T, n_local, info.obs_dim, device=dev),
                 actions=torch.empty(T, n_local, dtype=torch.int32, device=dev),
                 reward=torch.empty(T, n_local, device=dev),
                 done=torch.empty(T, n_local, dtype=torch.uint8, device=dev)) for _ in range(n_slots)]
    trajs = [_native.Traj(obs=r["obs"].data_ptr(), actions=r["actions"].data_ptr(), reward=r["reward"].data_ptr(),
                          done=r["done"].data_ptr()) for r in ring]
    lib, handle = env._lib, env._handle
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: fused rollout, device resident.
    # ONE pass = the K timed steps (len(plan) launches of the fused rollout kernel, each streaming its trajectory
    # to a ring slot; with N > 1 every launch also carries the obs all-gather). A pass of a few launches is
    # microseconds of GPU work, so the timed region is a TRAIN of R back-to-back passes (>= --min-gpu-seconds),
    # replayed from a CUDA graph of C passes so that the host's enqueue rate is not what is measured.
    launch_no = [0]

    def one_pass(st):
        for t_j in plan:
            j = launch_no[0]
            launch_no[0] += 1
            _native.check(lib.carlb_env_rollout(handle, t_j, 12345, (j * T) & 0x3FFFFFFF, None, _native.ACT_I32,
                                                ctypes.byref(trajs[j % n_slots]), st))
            if gather is not None and not fused_gather:
                gather.gather(lag=1 if j > 0 else 0)  # NCCL baseline: asynchronous collective, consumed one launch behind

    for w in range(max(3, -(-W // K))):  # >= W warm-up steps, eager
        one_pass(stream.cuda_stream)
    barrier()
    use_graph = gather is None or fused_gather
    passes_per_chunk = max(1, -(-n_slots // len(plan)))  # one trip round the ring
    while passes_per_chunk * len(plan) < 24:
        passes_per_chunk *= 2
    graph = None
    if use_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(passes_per_chunk):
                one_pass(torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)

    def chunk():
        if graph is not None:
            graph.replay()
        else:
            for _ in range(passes_per_chunk):
                one_pass(stream.cuda_stream)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _native.launch_count()
    train_ms, chunk_ms, n_chunks = timed_train(chunk, passes_per_chunk, stream, barrier, args.min_gpu_seconds)
    clock_info = clocks.stop() if rank == 0 else None
    if fused_gather:
        gather.resync()
    passes = n_chunks * passes_per_chunk
    launches = passes * len(plan)  # kernels of the timed region (graph replays are not seen by the host-side counter)
    train_ms = max_over_ranks(train_ms, dev, distributed)
    value = n_global * K * passes / (train_ms * 1e-3)
    pass_ms_median = float(np.median(chunk_ms)) / passes_per_chunk
    k_avg_ms = train_ms / launches  # launch-to-launch period inside the train (upper bound of the kernel duration)

    if args.fused_only:  # profiling runs: only the warm-up and the timed train, then stop
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                              "ms_per_step": train_ms / (passes * K), "fused_only": True, "gpu_launches": int(launches),
                              "config": {"fused_steps_per_launch": T, "launches_per_pass": len(plan), "passes": passes,
                                         "n_envs": n_global}, "kernel_ms_avg": k_avg_ms}))
        if distributed:
            dist.destroy_process_group()
        return 0

    # gathered tensor == what ONE GPU computes for the whole batch (correctness of the sharded path in this very run)
    gather_check = None
    if distributed:
        chk = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True, shard=(rank, world))
        from carl_b200.parallel import ObsGather

        cg = ObsGather(chk, mode=gather.mode, pipelined=False)
        chk.reset(seed=7)
        for j in range(3):
            chk.rollout(11, policy_seed=99, step_base=11 * j)
        got = cg.gather().clone()
        sha = tensor_sha(got)
        torch.cuda.synchronize(dev)
        want_sha = None
        if rank == 0:
            whole = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True)
            whole.reset(seed=7)
            for j in range(3):
                whole.rollout(11, policy_seed=99, step_base=11 * j)
            want_sha = tensor_sha(whole._obs)
            whole.close()
        shas = [None] * world
        dist.all_gather_object(shas, sha)
        if rank == 0:
            gather_check = {"equal": all(s_ == want_sha for s_ in shas), "sha256_16": want_sha, "ranks_checked": world,
                            "what": f"[{n_global}, {info.obs_dim}] gathered obs after reset(seed=7) + 3 x 11 fused steps on every "
                                    f"rank vs the same sequence on ONE GPU holding all {n_global} envs"}
        dist.barrier()
        cg.close()
        chk.close()

    # ---------------- step_api: one launch per step (+ gather with N > 1), graph-replayed, actions ring > L2
    ring_steps = int(math.ceil(1.2 * L2_BYTES / (n_local * 4)))  # int32 actions: 256 KiB per step
    G = ring_steps
    act_ring = torch.randint(0, 2, (G, n_local), dtype=torch.int32, device=dev)
    api_gather_mode = None
    if fused_gather:
        # the consumer that feeds obs k into step k+1 needs the gathered tensor of THIS step: SYNC mode
        _native.check(lib.carlb_gather_set_mode(gather._g, 0))
        api_gather_mode = "fused, sync (push + wait inside every step launch)"
    side2 = torch.cuda.Stream(dev)
    side2.wait_stream(stream)
    with torch.cuda.stream(side2):
        for g_ in range(3):
            _native.check(lib.carlb_env_step(handle, act_ring[g_ % G].data_ptr(), _native.ACT_I32, side2.cuda_stream))
    torch.cuda.synchronize(dev)
    api_graph = None
    if use_graph:
        api_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(api_graph, stream=side2):
            for g_ in range(G):
                _native.check(lib.carlb_env_step(handle, act_ring[g_].data_ptr(), _native.ACT_I32,
                                                 torch.cuda.current_stream(dev).cuda_stream))

    def api_replay():
        if api_graph is not None:
            api_graph.replay()
        else:
            for g_ in range(G):
                _native.check(lib.carlb_env_step(handle, act_ring[g_].data_ptr(), _native.ACT_I32, stream.cuda_stream))
                gather.gather()
            api_gather_mode_nccl[0] = "nccl all_gather_into_tensor after every step"

    api_gather_mode_nccl = [None]
    api_ms, _, api_reps = timed_train(api_replay, G, stream, barrier, min(args.min_gpu_seconds, 0.1), min_replays=2)
    api_ms = max_over_ranks(api_ms, dev, distributed)
    K_api = api_reps * G
    api_value = n_global * K_api / (api_ms * 1e-3)
    if fused_gather:
        gather.resync()
        _native.check(lib.carlb_gather_set_mode(gather._g, 1))

    # cold-L2 single launches: flush L2 (write a buffer larger than L2) between timed launches
    flush = torch.empty(L2_BYTES * 2 // 4, dtype=torch.float32, device=dev)
    cold = []
    for g_ in range(20):
        flush.fill_(float(g_))
        if distributed:
            dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        _native.check(lib.carlb_env_step(handle, act_ring[g_ % G].data_ptr(), _native.ACT_I32, stream.cuda_stream))
        c1.record(stream)
        torch.cuda.synchronize(dev)
        cold.append(c0.elapsed_time(c1))
    cold_ms = float(np.median(cold[3:]))
    del flush

    # ---------------- e2e: env.step(numpy actions) -> numpy results (host buffers, copies timed).
    # Passes of K calls, repeated until >= 2000 calls, per-pass times recorded.
    # (a policy writes into the same few page-locked buffers every step -- 4 here; cycling through many MB of
    # them only adds IOTLB misses on the GPU's PCIe reads: tools/e2e_probe.py, profiles/r01l_e2e_probe.json)
    E2E_ROWS = 4
    host_actions = hostmem.pinned_empty((E2E_ROWS, n_local), np.int32)
    host_actions[...] = np.random.default_rng(1).integers(0, 2, size=(E2E_ROWS, n_local), dtype=np.int32)
    # the host-facing API hands every rank's results to its own process: no cross-GPU gather on this path
    env_h = make_env() if distributed else env
    for w in range(max(5, min(W, 50))):
        env_h.step(host_actions[w % E2E_ROWS])
    e2e_passes = max(3, -(-args.e2e_calls // K))
    barrier()
    pass_s = []
    acc = 0.0
    t_begin = time.perf_counter()
    j = 0
    for p_ in range(e2e_passes):
        t0 = time.perf_counter()
        for _ in range(K):
            obs, rew, term, trunc, _i = env_h.step(host_actions[j % E2E_ROWS])
            acc += float(rew[0])  # the result is read on the host every step
            j += 1
        pass_s.append(time.perf_counter() - t0)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t_begin, dev, distributed)
    K_e2e = e2e_passes * K
    e2e_value = n_global * K_e2e / e2e_s
    h2d = n_local * 4
    d2h = n_local * (info.obs_dim * 4 + 4 + 1 + 1)
    e2e_async = e2e_async_leg(env_h, host_actions, K, e2e_passes, dev, distributed, n_global)

    peak, peak_src = hbm_peak()
    extra = {}
    if not args.no_ant:
        extra["ant_8192"] = ant_leg(dev, peak, args, rank, world, barrier)
        extra["config5_halfcheetah_hopper"] = config5_leg(dev, args, rank, world, barrier)
        extra["config3_pendulum_acrobot"] = config3_leg(dev, args, rank, world, barrier)
    if not args.no_f64:
        extra["value_f64"] = f64_leg(dev, names, table, rank, world, K, T, barrier, args, n_global)
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # DRAM bytes of ONE launch from the committed ncu --set full capture -- only quoted when that capture
    # fused the same number of env-steps per launch as this run (per launch, like `achieved`)
    prof, prof_file = measured_profile("prof_rollout", T)
    prof_traffic = float(prof["dram_bytes_per_launch"]) if "dram_bytes_per_launch" in prof else None
    prof_traffic_src = (f"profiles/{prof_file} (ncu --set full of `bench.py --steps {T} --fused-only`, {T} env-steps per launch; "
                        f"ncu flushes the caches before every replay and the 126 MB L2 still holds most of the launch's "
                        f"trajectory writes when the kernel ends)") if prof_file else None
    fused_bytes_per_launch = (TRAJ_BYTES * T + STEP_CONTRACT_BYTES) * n_local  # trajectory + one state/ctx round trip
    achieved = fused_bytes_per_launch / (k_avg_ms * 1e-3) / 1e9
    api_ms_per_launch = api_ms / K_api
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": train_ms / (passes * K), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "n_envs": n_global, "policy": "uniform random", "autoreset": True, "time_limit": 500,
            "policy_rng": "Philox4x32-10 in-kernel (GPU arm)",
            "fused_steps_per_launch": T, "launches_per_pass": len(plan),
            "timed_region": f"{passes} back-to-back passes of the {K}-step plan ({launches} rollout launches, "
                            f"{train_ms:.1f} ms of GPU time) between two CUDA events; "
                            + (f"replayed from a CUDA graph of {passes_per_chunk} passes" if graph is not None else "eager launches"),
            "passes": passes, "pass_ms_median": pass_ms_median, "pass_ms_mean": train_ms / passes,
            "l2": f"trajectory ring {n_slots} x {slot_bytes / 2**20:.0f} MiB > L2 (outputs go to DRAM); env state "
                  f"({n_local * 90 / 2**20:.1f} MiB working set) is register/L2 resident by design",
            "collective": (f"obs all-gather with EVERY rollout launch: {gather_mode}"
                           + (f" ({gather.transport}; pipelined: a publisher warp per CTA pushes obs k over NVLink while launch "
                              f"k+1 computes, flags and waits in-kernel, no extra launch)" if fused_gather else
                              " (asynchronous NCCL collective, consumed one launch behind)")) if distributed else "none",
        },
        "gpu_launches": int(launches),
        "gpu_launches_host_counted": int(_native.launch_count() - launches0),
        "clocks": clock_info,
        "e2e": None,
        "e2e_sync": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "steps": K_e2e, "passes": e2e_passes, "pass_ms_median": float(np.median(pass_s)) * 1e3,
                     "api": "CARLCartPole.step(numpy int32 actions in page-locked memory) -> numpy obs/reward/terminated/"
                            "truncated, one synchronous call per step (carlb_env_step_host_checked: the step kernel reads the "
                            "actions over PCIe, range-checks them itself -- undo log, rolled back if one is invalid -- writes the "
                            "results over PCIe and stores a completion word the call polls)",
                     "ms_per_step": e2e_s / K_e2e * 1e3},
        "roofline": {
            "kernel": "rollout_kernel<CARTPOLE,float> (fused T-step rollout, trajectory to HBM)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": prof_traffic,
            "traffic_source": prof_traffic_src,
            "algorithmic_bytes_per_launch": fused_bytes_per_launch,
            "bytes_per_env_step": TRAJ_BYTES + STEP_CONTRACT_BYTES / T,
            "kernel_ms_avg": k_avg_ms,
            "kernel_ms_how": "train time / launches: the launch-to-launch period inside the timed train (includes the "
                             "inter-kernel gap, so it bounds the kernel duration from above)",
            "frac_vs_step_contract_90B": (STEP_CONTRACT_BYTES * n_local * T / (k_avg_ms * 1e-3) / 1e9) / peak,
            "frac_note": "`frac` counts the bytes the fused kernel must move (25 B trajectory per env-step + one 90 B "
                         "state/context round trip per launch); SURVEY §8(d)'s per-step contract (90 B per env-step, state "
                         "re-read every step) gives frac_vs_step_contract_90B, which can exceed 1 because the state stays in registers",
            # the kernel is instruction-issue bound, not HBM bound, at this batch size (ncu, same T):
            "issue_slot_utilisation_pct": prof.get("issue_active_pct"),
            "warps_active_pct": prof.get("warps_active_pct"),
            "ncu_kernel_us": prof.get("duration_us_under_ncu"),
        },
        "step_api": {
            "value": api_value, "unit": UNIT, "steps": K_api, "us_per_launch": api_ms_per_launch * 1e3,
            "l2": f"actions ring of {G} steps x 256 KiB > L2; " + (f"CUDA-graph replay of {G} launches" if api_graph is not None else "eager launches"),
            "gather": (api_gather_mode or api_gather_mode_nccl[0]) if distributed else None,
            "roofline": {"kernel": "step_kernel<CARTPOLE,float>", "bound": "hbm",
                         "achieved": STEP_CONTRACT_BYTES * n_local / (api_ms_per_launch * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": STEP_CONTRACT_BYTES * n_local / (api_ms_per_launch * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": STEP_CONTRACT_BYTES * n_local},
            "cold_l2_us_per_launch": cold_ms * 1e3,
            "cold_l2_frac": STEP_CONTRACT_BYTES * n_local / (cold_ms * 1e-3) / 1e9 / peak,
        },
    }
    if e2e_async is not None:
        # the headline end-to-end number: the same 65 536 envs per GPU stepped through the split-batch host API
        out["e2e"] = dict(e2e_async, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                          api="CARLCartPole.step_async(numpy int32 actions of one half of the batch, page-locked) / "
                              "step_wait(part) -> numpy obs/reward/terminated/truncated of that half, the two halves alternating "
                              "(EnvPool-style split-batch stepping; carlb_env_step_host_begin / _end: in-kernel action check with "
                              "undo log, results written over PCIe by the kernel, completion word polled by the host). Every step "
                              "all envs advance once, all actions cross PCIe host->device and all results device->host")
    else:
        out["e2e"] = out["e2e_sync"]
    if gather_check is not None:
        out["gather_check"] = gather_check
    out.update(extra)
    # what the parity claims rest on; "unpinned" = no reference-run vector exists or can be generated here
    out["parity"] = {
        "pinned_by_published_known_answers": ["CARLCartPole step + reset (gymnasium known answers)", "ContextSampler stream "
                                              "(reference notebooks)", "PCG64 / SeedSequence reset streams (numpy)",
                                              "Philox4x32-10, threefry2x32 (Random123 KATs)",
                                              "Brax reset stream: JAX PRNGKey / split / uniform / normal (JAX documentation outputs)"],
        "anchored_to_independent_physics": ["CARLPendulum (uniform-rod period, torque response, O(dt) energy drift)",
                                            "CARLAcrobot (one step == RK4 of the textbook manipulator equations, 1e-10)",
                                            "CARLMountainCar / Continuous (valley period of the symplectic map)",
                                            "Brax spring pipeline: free fall closed form (dt x n_frames, gravity), momentum "
                                            "conservation, cart-pendulum frequency; geometry by CARL's own mass defaults"],
        "unpinned": ["Pendulum / Acrobot / MountainCar conventions no textbook fixes (clip order, reward constants): from SURVEY App. A",
                     "all seven Brax bodies against real brax 0.12.1 (not installable: profiles/r02a_pip_install_attempt.txt); "
                     "kernel-vs-oracle parity is exact to the stated tolerances, oracle-vs-Brax is not verifiable offline"],
        "gpu_tests": "tests -m gpu: CUDA through the C ABI vs the CPU oracle (profiles/r02o_pytest_gpu.txt)",
    }
    done_counts = committed_done_mask_counts()
    if done_counts is not None:
        out["done_mask_mismatches_fp32"] = done_counts
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_leg(steps_per_call=200, target_seconds=args.cpu_seconds)
    print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()
    return 0


def committed_done_mask_counts():
    """fp32-mode done-mask mismatches against the float64 oracle, counted by the GPU test-suite on 10^6 random
    (state, action, context) triples per env kind (tests/test_classic_parity_gpu.py writes the file; bench.py
    itself never runs the oracle outside its CPU-baseline leg)."""
    p = os.path.join(ROOT, "profiles", "done_mask_counts.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def e2e_async_leg(env, host_actions, K, passes, dev, distributed, n_global, parts=2):
    """The split-batch host-buffer API (EnvPool send / recv, SB3 step_async / step_wait): the batch is stepped as
    `parts` contiguous parts on their own streams; while the host reads part p's results and hands in its next
    actions, the other part's results are crossing PCIe. One "step" = every part stepped once."""
    import torch

    env.async_parts = parts
    rows = host_actions.shape[0]
    bounds = [env.part_range(p) for p in range(parts)]
    for w in range(10):
        env.step_async(host_actions[w % rows])
        env.step_wait()
    if distributed:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize(dev)
    for p, (lo, hi) in enumerate(bounds):
        env.step_async(host_actions[0, lo:hi], part=p)
    acc = 0.0
    pass_s = []
    j = 0
    t_begin = time.perf_counter()
    for _ in range(passes):
        t0 = time.perf_counter()
        for _k in range(K):
            row = host_actions[(j + 1) % rows]
            for p, (lo, hi) in enumerate(bounds):
                st, rew, term, trunc, _i = env.step_wait(part=p)
                acc += float(rew[0])  # the part's result is read on the host every step
                env.step_async(row[lo:hi], part=p)
            j += 1
        pass_s.append(time.perf_counter() - t0)
    e2e_s = max_over_ranks(time.perf_counter() - t_begin, dev, distributed)
    for p in range(parts):
        env.step_wait(part=p)
    steps = passes * K
    return {"value": n_global * steps / e2e_s, "unit": UNIT, "steps": steps, "passes": passes, "parts": parts,
            "ms_per_step": e2e_s / steps * 1e3, "pass_ms_median": float(np.median(pass_s)) * 1e3}


def f64_leg(dev, names, table, rank, world, K, T, barrier, args, n_global):
    """The same fused-rollout train in the reference's own precision (dtype="float64": state, context and
    arithmetic in float64, done masks bit-identical to the float64 reference)."""
    import ctypes

    import torch

    from carl_b200 import _native
    from carl_b200.envs import CARLCartPole, ContextTable

    n_local = N_ENVS_PER_GPU
    env = CARLCartPole(contexts=ContextTable(names, table), device=dev, autoreset=True, shard=(rank, world), dtype="float64")
    env.reset(seed=0)
    slot_bytes = T * n_local * TRAJ_BYTES
    n_slots = max(2, int(math.ceil(1.5 * L2_BYTES / slot_bytes)) + 1)
    ring = [dict(obs=torch.empty(T, n_local, 4, device=dev), actions=torch.empty(T, n_local, dtype=torch.int32, device=dev),
                 reward=torch.empty(T, n_local, device=dev), done=torch.empty(T, n_local, dtype=torch.uint8, device=dev))
            for _ in range(n_slots)]
    trajs = [_native.Traj(obs=r["obs"].data_ptr(), actions=r["actions"].data_ptr(), reward=r["reward"].data_ptr(),
                          done=r["done"].data_ptr()) for r in ring]
    stream = torch.cuda.current_stream(dev)
    plan = [T] * (K // T) + ([K % T] if K % T else [])
    cnt = [0]

    def one_pass(st):
        for t_j in plan:
            j = cnt[0]
            cnt[0] += 1
            _native.check(env._lib.carlb_env_rollout(env._handle, t_j, 12345, (j * T) & 0x3FFFFFFF, None, _native.ACT_I32,
                                                     ctypes.byref(trajs[j % n_slots]), st))

    for _ in range(3):
        one_pass(stream.cuda_stream)
    torch.cuda.synchronize(dev)
    ppc = max(1, -(-n_slots // len(plan)))
    side = torch.cuda.Stream(dev)
    side.wait_stream(stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(ppc):
            one_pass(torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    ms, _, reps = timed_train(graph.replay, ppc, stream, barrier, min(args.min_gpu_seconds, 0.1), min_replays=3)
    ms = max_over_ranks(ms, dev, world > 1)
    env.close()
    return {"value": n_global * K * ppc * reps / (ms * 1e-3), "unit": UNIT, "dtype": "f64", "passes": ppc * reps,
            "note": "same workload and plan as `value`, CARLCartPole(dtype='float64'): the reference's float64 arithmetic"}


def _brax_contexts(cls, n, features):
    from carl_b200.context import ContextSampler, UniformFloatContextFeature
    from carl_b200.envs import ContextTable

    names = list(cls.get_context_space().get_default_context().keys())
    sampler = ContextSampler([UniformFloatContextFeature(k, lo, hi) for k, (lo, hi) in features.items()],
                             context_space=cls.get_context_space(), seed=0)
    return ContextTable(names, sampler.sample_context_table(n, names))


def ant_leg(dev, peak, args, rank=0, world=1, barrier=None):
    """Secondary north-star workload (BASELINE configs[3]): CARLBraxAnt, 8 192 contexts PER GPU (weak scaling over
    1 -> 8 GPUs; gravity / mass_torso / friction sampled, context_mode="applied"), uniform random policy. With N > 1
    every launch carries the obs all-gather (fused NVLink push, pipelined)."""
    import ctypes
    import time

    import torch

    from carl_b200 import _native, hostmem
    from carl_b200.envs import CARLBraxAnt

    distributed = world > 1
    n = 8192
    n_global = n * world
    ctxs = _brax_contexts(CARLBraxAnt, n_global, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)})
    env = CARLBraxAnt(contexts=ctxs, device=dev, context_mode="applied", shard=(rank, world))
    env.reset(seed=0)
    gather = None
    if distributed:
        from carl_b200.parallel import ObsGather

        gather = ObsGather(env, mode="fused", pipelined=True)
    info = env._info
    T = 20
    traj_bytes = info.obs_dim * 4 + info.act_dim * 4 + 4 + 1
    ring = [dict(obs=torch.empty(T, n, info.obs_dim, device=dev), actions=torch.empty(T, n, info.act_dim, device=dev),
                 reward=torch.empty(T, n, device=dev), done=torch.empty(T, n, dtype=torch.uint8, device=dev))
            for _ in range(8)]  # 8 x 23 MiB > L2
    trajs = [_native.Traj(obs=r["obs"].data_ptr(), actions=r["actions"].data_ptr(), reward=r["reward"].data_ptr(),
                          done=r["done"].data_ptr()) for r in ring]
    stream = torch.cuda.current_stream(dev)
    cnt = [0]

    def launch():
        j = cnt[0]
        cnt[0] += 1
        _native.check(env._lib.carlb_env_rollout(env._handle, T, 7, j * T, None, _native.ACT_F32,
                                                 ctypes.byref(trajs[j % 8]), stream.cuda_stream))

    for j in range(3):
        launch()
    ms, per, reps = timed_train(launch, 1, stream, barrier, min(args.min_gpu_seconds, 0.15), min_replays=5)
    ms = max_over_ranks(ms, dev, distributed)
    fused = n_global * T * reps / (ms * 1e-3)
    launch_ms = ms / reps
    # the same train with the FMA-contracted build of the kernels (arithmetic="fma", carlb_brax_set_arithmetic)
    _native.check(env._lib.carlb_brax_set_arithmetic(env._handle, 1))
    for j in range(2):
        launch()
    ms_f, _, reps_f = timed_train(launch, 1, stream, barrier, min(args.min_gpu_seconds, 0.1), min_replays=5)
    fused_fma = n_global * T * reps_f / (max_over_ranks(ms_f, dev, distributed) * 1e-3)
    _native.check(env._lib.carlb_brax_set_arithmetic(env._handle, 0))
    # single-step API, device actions (+ the gather of every step's obs in SYNC mode with N > 1)
    if gather is not None:
        _native.check(env._lib.carlb_gather_set_mode(gather._g, 0))
    acts = torch.rand(64, n, info.act_dim, device=dev) * 2 - 1
    for j in range(5):
        env.step(acts[j])
    cnt2 = [0]

    def step_launch():
        j = cnt2[0]
        cnt2[0] += 1
        _native.check(env._lib.carlb_env_step(env._handle, acts[j % 64].data_ptr(), _native.ACT_F32, stream.cuda_stream))

    api_total_ms, _, api_reps = timed_train(step_launch, 1, stream, barrier, 0.03, min_replays=50)
    api_ms = max_over_ranks(api_total_ms, dev, distributed) / api_reps
    step_bytes = 1114  # R state 468 + ctx 24 + action 32 + 4 ; W state 468 + obs 108 + 4 + 2 + 4
    # host buffers in / out: env.step(numpy float32 actions in page-locked memory) -> numpy results
    host_acts = hostmem.pinned_empty((4, n, info.act_dim), np.float32)
    host_acts[...] = np.random.default_rng(2).uniform(-1, 1, size=host_acts.shape).astype(np.float32)
    for j in range(5):
        env.step(host_acts[j % 4])
    barrier()
    t0 = time.perf_counter()
    acc = 0.0
    n_e2e = 200
    for j in range(n_e2e):
        o_h, r_h, te_h, tr_h, _ = env.step(host_acts[j % 4])
        acc += float(r_h[0])
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks(time.perf_counter() - t0, dev, distributed) / n_e2e * 1e3
    # the same host-buffer loop with the FMA build (arithmetic="fma"): the shorter kernel shortens the serial part
    _native.check(env._lib.carlb_brax_set_arithmetic(env._handle, 1))
    for j in range(5):
        env.step(host_acts[j % 4])
    barrier()
    t0 = time.perf_counter()
    for j in range(n_e2e):
        o_h, r_h, te_h, tr_h, _ = env.step(host_acts[j % 4])
        acc += float(r_h[0])
    torch.cuda.synchronize(dev)
    e2e_fma_ms = max_over_ranks(time.perf_counter() - t0, dev, distributed) / n_e2e * 1e3
    _native.check(env._lib.carlb_brax_set_arithmetic(env._handle, 0))
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:  # reported next to the GPU number; never allowed to break the bench line
            cpu = brax_cpu_baseline(env._sysd, env._ctx.cpu().numpy(), target_seconds=3.0)
        except Exception as e:  # pragma: no cover
            cpu = {"error": repr(e)}
    transport = gather.transport if gather is not None else None
    if gather is not None:
        torch.cuda.synchronize(dev)
        barrier()
        gather.close()
    env.close()
    return {
        "cpu_baseline": cpu,
        "workload": f"CARLBraxAnt, 8192 sampled contexts per GPU (gravity/mass_torso/friction), context_mode=applied, "
                    f"uniform random policy, 10 spring substeps per env-step; {n_global} envs on {world} GPU(s), weak scaling",
        "value": fused, "unit": UNIT, "fused_steps_per_launch": T, "launches": reps, "ms_per_env_step_batch": launch_ms / T,
        "value_fma": fused_fma,
        "arithmetic": "value: strict (products and sums rounded separately; the parity mode); value_fma: FMA-contracted build "
                      "(arithmetic='fma'; held to the float64 yardstick by tests/test_brax_parity_gpu.py)",
        "collective": (f"obs all-gather with every launch: fused ({transport}), pipelined" if distributed else "none"),
        "roofline": {"bound": "fp32-issue (HBM shown for reference)", "bytes_per_env_step": traj_bytes + step_bytes / T,
                     "achieved": (traj_bytes * T + step_bytes) * n / (launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": (traj_bytes * T + step_bytes) * n / (launch_ms * 1e-3) / 1e9 / peak,
                     "frac_vs_step_contract_1114B": step_bytes * n * T / (launch_ms * 1e-3) / 1e9 / peak},
        # the bound that matters for this kernel: instruction issue (per GPU; ncu instruction counts of the same launch)
        "issue_roofline": issue_roof("prof_brax", n * T, fused / world, None),
        "issue_roofline_fma": issue_roof("prof_brax_fma", n * T, fused_fma / world, None),
        "step_api": {"value": n_global / (api_ms * 1e-3), "us_per_launch": api_ms * 1e3,
                     "gather": "fused, sync (push + wait inside every step launch)" if distributed else None,
                     "hbm_frac_1114B": step_bytes * n / (api_ms * 1e-3) / 1e9 / peak},
        "e2e": {"value": n_global / (e2e_ms * 1e-3), "unit": UNIT, "us_per_step": e2e_ms * 1e3,
                "value_fma": n_global / (e2e_fma_ms * 1e-3), "us_per_step_fma": e2e_fma_ms * 1e3,
                "h2d_bytes_per_step": n * info.act_dim * 4, "d2h_bytes_per_step": n * (info.obs_dim * 4 + 4 + 1 + 1),
                "api": "CARLBraxAnt.step(numpy float32 actions in page-locked memory) -> numpy obs/reward/terminated/truncated "
                       "(the step kernel reads the actions and writes the results over PCIe itself)"},
    }


def config3_leg(dev, args, rank, world, barrier):
    """BASELINE configs[2]: CARLPendulum + CARLAcrobot, 32 768 contexts each PER GPU, advanced by ONE mixed launch
    per step (`carlb_mixed_step`: block-uniform kind switch), replayed from a CUDA graph; plus the two shards as fused
    rollouts (20 steps per launch)."""
    import ctypes

    import torch

    from carl_b200 import _native
    from carl_b200.context import ContextSampler, UniformFloatContextFeature
    from carl_b200.envs import CARLAcrobot, CARLPendulum, ContextTable, MixedBatch

    n = 32768

    def table(cls, feats):
        names = list(cls.get_context_space().get_default_context().keys())
        smp = ContextSampler([UniformFloatContextFeature(k, lo, hi) for k, (lo, hi) in feats.items()],
                             context_space=cls.get_context_space(), seed=0)
        return ContextTable(names, smp.sample_context_table(n * world, names))

    pen = CARLPendulum(contexts=table(CARLPendulum, {"g": (5, 15), "m": (0.5, 2), "l": (0.5, 2)}), device=dev, autoreset=True,
                       shard=(rank, world))
    acr = CARLAcrobot(contexts=table(CARLAcrobot, {"LINK_MASS_1": (0.5, 2), "LINK_MASS_2": (0.5, 2), "LINK_LENGTH_1": (0.5, 2)}),
                      device=dev, autoreset=True, shard=(rank, world))
    mixed = MixedBatch([pen, acr])
    mixed.reset(seed=0)
    G = 64
    ap = torch.rand(G, n, device=dev) * 4 - 2
    aa = torch.randint(0, 3, (G, n), dtype=torch.int32, device=dev)
    dts = (ctypes.c_int * 2)(_native.ACT_F32, _native.ACT_I32)
    stream = torch.cuda.current_stream(dev)

    def step(j, st):
        ptrs = (ctypes.c_void_p * 2)(ap[j % G].data_ptr(), aa[j % G].data_ptr())
        _native.check(mixed._lib.carlb_mixed_step(mixed._handles, ptrs, dts, 2, st))

    for j in range(3):
        step(j, stream.cuda_stream)
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(dev)
    side.wait_stream(stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for j in range(G):
            step(j, torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    ms, _, reps = timed_train(graph.replay, G, stream, barrier, min(args.min_gpu_seconds, 0.05), min_replays=3)
    ms = max_over_ranks(ms, dev, world > 1)
    us = ms / (reps * G) * 1e3
    out = {"workload": f"CARLPendulum {n} + CARLAcrobot {n} contexts per GPU, ONE mixed launch per step (CUDA-graph replay), "
                       f"{world} GPU(s)",
           "value": 2 * n * world / (us * 1e-6), "unit": UNIT, "us_per_step": us,
           "algorithmic_GBps_62B_110B": (62 + 110) * n / (us * 1e-6) / 1e9}
    T = 20
    for env, name in ((pen, "pendulum"), (acr, "acrobot")):
        fn = lambda: env.rollout(T, policy_seed=1, record=False)
        for _ in range(2):
            fn()
        ms_r, _, reps_r = timed_train(fn, 1, stream, barrier, 0.03, min_replays=5)
        ms_r = max_over_ranks(ms_r, dev, world > 1)
        out[f"{name}_fused_{T}_steps"] = {"value": n * world * T * reps_r / (ms_r * 1e-3), "unit": UNIT}
    pen.close()
    acr.close()
    return out


def config5_leg(dev, args, rank, world, barrier):
    """BASELINE configs[4]: CARLBraxHalfcheetah + CARLBraxHopper, 16 384 contexts in total (8 192 each), sharded over
    the N GPUs (STRONG scaling: the total is fixed), one `step` per env kind per step with the obs all-gather of
    EVERY step (fused NVLink push, sync mode: the gathered tensor of step k is complete when step k's launch ends).
    The two kinds run on two streams (independent batches)."""
    import torch

    from carl_b200 import _native
    from carl_b200.envs import CARLBraxHalfcheetah, CARLBraxHopper

    distributed = world > 1
    n_each = 8192
    envs, gathers, acts, streams = [], [], [], []
    for cls in (CARLBraxHalfcheetah, CARLBraxHopper):
        ctxs = _brax_contexts(cls, n_each, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)})
        e = cls(contexts=ctxs, device=dev, context_mode="applied", shard=(rank, world))
        e.reset(seed=0)
        envs.append(e)
        if distributed:
            from carl_b200.parallel import ObsGather

            gathers.append(ObsGather(e, mode="fused", pipelined=False))
        acts.append(torch.rand(16, e.num_envs, e._info.act_dim, device=dev) * 2 - 1)
        streams.append(torch.cuda.Stream(dev))
    main = torch.cuda.current_stream(dev)
    cnt = [0]
    fork, joins = torch.cuda.Event(), [torch.cuda.Event() for _ in envs]

    def step_both():
        j = cnt[0]
        cnt[0] += 1
        fork.record(main)
        for e, a, st, jn in zip(envs, acts, streams, joins):
            st.wait_event(fork)
            _native.check(e._lib.carlb_env_step(e._handle, a[j % 16].data_ptr(), _native.ACT_F32, st.cuda_stream))
            jn.record(st)
            main.wait_event(jn)

    for _ in range(5):
        step_both()
    ms, _, reps = timed_train(step_both, 1, main, barrier, min(args.min_gpu_seconds, 0.1), min_replays=20)
    ms = max_over_ranks(ms, dev, distributed)
    transport = gathers[0].transport if gathers else None
    torch.cuda.synchronize(dev)
    barrier()
    for g in gathers:
        g.close()
    for e in envs:
        e.close()
    return {"workload": f"CARLBraxHalfcheetah 8192 + CARLBraxHopper 8192 contexts in total, sharded over {world} GPU(s) "
                        f"({n_each // world} of each per GPU), one step per kind per step, obs all-gather EVERY step",
            "value": 2 * n_each * reps / (ms * 1e-3), "unit": UNIT, "scaling": "strong", "steps": reps,
            "us_per_step": ms / reps * 1e3,
            "collective": f"fused ({transport}), sync: push + wait inside every step launch" if distributed else "none"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="carl_b200", choices=["carl_b200", "reference"])
    ap.add_argument("--fuse", type=int, default=500, help="env-steps fused per rollout launch")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ant", action="store_true", help="skip the CARLBraxAnt and Halfcheetah+Hopper legs")
    ap.add_argument("--no-f64", action="store_true", help="skip the float64 leg")
    ap.add_argument("--min-gpu-seconds", type=float, default=0.25,
                    help="the timed train repeats the K-step pass until at least this much GPU time is covered")
    ap.add_argument("--e2e-calls", type=int, default=2000, help="minimum number of timed env.step(numpy) calls")
    ap.add_argument("--fused-only", action="store_true",
                    help="profiling aid: run the warm-up and the timed fused-rollout region, print a short line, stop")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="multi-GPU obs gather path")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
