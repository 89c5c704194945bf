"""Secondary BASELINE configs on one GPU (numbers for DESIGN.md / profiles; not the headline line):
config 3 -- CARLPendulum + CARLAcrobot, 32 768 contexts each, ONE mixed launch per step;
config 5 (single-GPU share) -- CARLBraxHalfcheetah + CARLBraxHopper, 8 192 contexts each."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from carl_b200.context import ContextSampler, UniformFloatContextFeature
from carl_b200.envs import (CARLAcrobot, CARLBraxHalfcheetah, CARLBraxHopper, CARLBraxHumanoid, CARLBraxHumanoidStandup, CARLBraxPusher,
                            CARLBraxInvertedDoublePendulum, CARLBraxInvertedPendulum, CARLBraxReacher, CARLBraxWalker2d,
                            CARLPendulum, ContextTable, MixedBatch)


def table(cls, feats, n):
    names = list(cls.get_context_space().get_default_context().keys())
    s = ContextSampler([UniformFloatContextFeature(k, lo, hi) for k, (lo, hi) in feats.items()],
                       context_space=cls.get_context_space(), seed=0)
    return ContextTable(names, s.sample_context_table(n, names))


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


def brax_bodies(out, bodies):
    nb = 8192
    for cls, feats, name in bodies:
        for arithmetic in ("strict", "fma"):
            env = cls(contexts=table(cls, feats, nb), context_mode="applied", arithmetic=arithmetic)
            env.reset(seed=0)
            T = 20
            ms = timed(lambda: env.rollout(T, policy_seed=1, record=False), 10, warm=2)
            a = (torch.rand(nb, env._info.act_dim, device="cuda") * 2 - 1) * env._sysd["act_scale"]
            ms1 = timed(lambda: env.step(a), 50)
            key = f"config5_{name}" + ("" if arithmetic == "strict" else "_fma")
            out[key] = {"workload": f"{cls.__name__} {nb} contexts", "fused_env_steps_per_s": nb * T / (ms * 1e-3),
                        "step_api_env_steps_per_s": nb / (ms1 * 1e-3), "step_us": ms1 * 1e3}


HUMANOIDS = ((CARLBraxHumanoid, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}, "humanoid"),
             (CARLBraxHumanoidStandup, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}, "humanoidstandup"),
             (CARLBraxPusher, {"gravity": (-1e-3, -1e-6), "mass_object": (1e-3, 3e-3), "friction": (0.5, 1.5)}, "pusher"))


def main():
    out = {}
    if "--only-humanoid" in sys.argv:
        brax_bodies(out, HUMANOIDS)
        print(json.dumps(out))
        return
    n = 32768
    pen = CARLPendulum(contexts=table(CARLPendulum, {"g": (5, 15), "m": (0.5, 2), "l": (0.5, 2)}, n), autoreset=True)
    acr = CARLAcrobot(contexts=table(CARLAcrobot, {"LINK_MASS_1": (0.5, 2), "LINK_MASS_2": (0.5, 2), "LINK_LENGTH_1": (0.5, 2)}, n),
                      autoreset=True)
    mixed = MixedBatch([pen, acr])
    mixed.reset(seed=0)
    ap = torch.rand(n, device="cuda") * 4 - 2
    aa = torch.randint(0, 3, (n,), dtype=torch.int32, device="cuda")
    ms = timed(lambda: mixed.step([ap, aa]), 500)
    out["config3_mixed_step"] = {"workload": "CARLPendulum 32768 + CARLAcrobot 32768, one mixed launch per step (Python API call)",
                                 "us_per_step": ms * 1e3, "env_steps_per_s": 2 * n / (ms * 1e-3),
                                 "algorithmic_GBps": (62 + 110) * n / (ms * 1e-3) / 1e9}
    # the same launch through the bare C ABI (argument arrays built once): the Python call above is host bound
    import ctypes

    from carl_b200 import _native

    ptrs = (ctypes.c_void_p * 2)(ap.data_ptr(), aa.data_ptr())
    dts = (ctypes.c_int * 2)(_native.ACT_F32, _native.ACT_I32)
    st = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: mixed._lib.carlb_mixed_step(mixed._handles, ptrs, dts, 2, st), 2000)
    out["config3_mixed_step_c_abi"] = {"us_per_step": ms * 1e3, "env_steps_per_s": 2 * n / (ms * 1e-3),
                                       "algorithmic_GBps": (62 + 110) * n / (ms * 1e-3) / 1e9}
    for env, name in ((pen, "pendulum"), (acr, "acrobot")):
        T = 100
        ms = timed(lambda: env.rollout(T, policy_seed=1, record=False), 10, warm=2)
        out[f"config3_{name}_fused"] = {"env_steps_per_s": n * T / (ms * 1e-3), "ms_per_100_steps": ms}
    brax_bodies(out, ((CARLBraxHalfcheetah, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}, "halfcheetah"),
                      (CARLBraxHopper, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}, "hopper"),
                      (CARLBraxWalker2d, {"gravity": (-15, -5), "mass_torso": (5, 20), "friction": (0.5, 1.5)}, "walker2d"),
                      (CARLBraxInvertedPendulum, {"gravity": (-15, -5), "mass_cart": (5, 20), "mass_pole": (2, 8)}, "inverted_pendulum"),
                      (CARLBraxInvertedDoublePendulum, {"gravity": (-15, -5), "mass_cart": (5, 20), "mass_pole": (2, 8)},
                       "inverted_double_pendulum"),
                      (CARLBraxReacher, {"gravity": (-15, -5), "mass_body0": (0.02, 0.06), "mass_body1": (0.02, 0.06)}, "reacher"))
                + HUMANOIDS)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
