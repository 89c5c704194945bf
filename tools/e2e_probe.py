"""Where does the host-buffer step (bench.py's `e2e`) spend its time?  Prints one JSON object.

Times, for CARLCartPole x 65 536 contexts: the Python API call, the bare C-ABI call on page-locked
buffers under the three CARLB_ZEROCOPY modes, int32 vs uint8 actions, the device-resident step with a
stream sync per step (launch + sync floor), and the native staging pass alone."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carl_b200 import _native  # noqa: E402
from carl_b200.context import ContextSampler, UniformFloatContextFeature  # noqa: E402
from carl_b200.envs import CARLCartPole, ContextTable  # noqa: E402


def loop(fn, n=400, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    n = int(os.environ.get("PROBE_N", 65536))
    names = list(CARLCartPole.get_context_space().get_default_context().keys())
    sampler = ContextSampler([UniformFloatContextFeature("gravity", 5, 15), UniformFloatContextFeature("length", 0.3, 0.8)],
                             context_space=CARLCartPole.get_context_space(), seed=0)
    env = CARLCartPole(contexts=ContextTable(names, sampler.sample_context_table(n, names)))
    env.reset(seed=0)
    lib, h = env._lib, env._handle
    io = env._ensure_host_io()
    p = io["ptrs"]
    st = env._stream()
    rng = np.random.default_rng(1)
    a32 = rng.integers(0, 2, size=(64, n), dtype=np.int32)
    a8 = a32.astype(np.uint8)
    out = {"n_envs": n}
    k = [0]

    def py_step():
        k[0] += 1
        o, r, te, tr, _ = env.step(a32[k[0] % 64])
        return float(r[0])

    out["python_env_step_int32_us"] = loop(py_step)
    from carl_b200 import hostmem
    for rows in (64, 4, 1):
        pa = hostmem.pinned_empty((rows, n), np.int32)
        pa[...] = a32[:rows]
        kk = [0]

        def pinned_step():
            kk[0] += 1
            o, r, te, tr, _ = env.step(pa[kk[0] % rows])
            return float(r[0])

        out[f"python_env_step_pinned_inplace_{rows}rows_us"] = loop(pinned_step)
        out[f"c_abi_step_host_pinned_inplace_{rows}rows_us"] = loop(
            lambda: lib.carlb_env_step_host(h, pa[kk[0] % rows].ctypes.data, _native.ACT_I32, p[1], p[2], p[3], p[4], st))
        out[f"c_abi_step_host_checked_pinned_inplace_{rows}rows_us"] = loop(
            lambda: lib.carlb_env_step_host_checked(h, pa[kk[0] % rows].ctypes.data, _native.ACT_I32, 2, p[1], p[2], p[3], p[4], st))
        out[f"check_only_{rows}rows_us"] = loop(lambda: lib.carlb_stage_actions(None, pa[0].ctypes.data, n, _native.ACT_I32, 2), n=2000)
        hostmem.release(pa)
    out["python_env_step_uint8_us"] = loop(lambda: env.step(a8[0]))
    out["stage_actions_int32_us"] = loop(lambda: lib.carlb_stage_actions(p[0], a32[3].ctypes.data, n, _native.ACT_I32, 2), n=2000)
    out["stage_actions_uint8_us"] = loop(lambda: lib.carlb_stage_actions(p[0], a8[3].ctypes.data, n, _native.ACT_U8, 2), n=2000)
    for mode in ("1", "2", "0"):
        os.environ["CARLB_ZEROCOPY"] = mode
        lib.carlb_stage_actions(p[0], a32[3].ctypes.data, n, _native.ACT_I32, 2)
        out[f"c_abi_step_host_zc{mode}_int32_us"] = loop(
            lambda: lib.carlb_env_step_host(h, p[0], _native.ACT_I32, p[1], p[2], p[3], p[4], st))
        lib.carlb_stage_actions(p[0], a8[3].ctypes.data, n, _native.ACT_U8, 2)
        out[f"c_abi_step_host_zc{mode}_uint8_us"] = loop(
            lambda: lib.carlb_env_step_host(h, p[0], _native.ACT_U8, p[1], p[2], p[3], p[4], st))
    os.environ["CARLB_ZEROCOPY"] = "1"
    dact = torch.from_numpy(a32[0]).cuda()

    def dev_step_sync():
        lib.carlb_env_step(h, dact.data_ptr(), _native.ACT_I32, st)
        torch.cuda.synchronize()

    out["device_step_plus_sync_us"] = loop(dev_step_sync)
    # raw PCIe: one pinned D2H copy of the result block / one H2D copy of the actions, synchronised
    dev_out = torch.empty(io["out"].numel(), dtype=torch.uint8, device="cuda")
    out["d2h_result_block_copy_sync_us"] = loop(lambda: (io["out"].copy_(dev_out, non_blocking=True), torch.cuda.synchronize()))
    dev_act = torch.empty(n, dtype=torch.int32, device="cuda")
    hact = io["act"][: n // 2].view(torch.int32)
    out["h2d_actions_copy_sync_us"] = loop(lambda: (dev_act.copy_(hact, non_blocking=True), torch.cuda.synchronize()))
    out["result_block_bytes"] = int(io["out"].numel())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
