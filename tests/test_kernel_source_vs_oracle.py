"""No-GPU check of the kernel *source logic*: the product's __host__ __device__ physics
(carl_b200/csrc/physics_classic.h), compiled by g++ through tests/hostcheck, against the CPU
oracle on the same seeded inputs. The same comparisons run against the real CUDA build in
tests/test_classic_parity_gpu.py."""
import numpy as np
import pytest

from oracle.classic import KINDS, OracleClassicEnv
from tests.util import HostCheck, done_margin, kernel_rows, sample_actions, sample_context_table, sample_states

KIND_LIST = list(KINDS)


@pytest.fixture(scope="module")
def hc():
    return HostCheck()


def _single_step(hc, kind, f64, n=2000, seed=0, mode="reference"):
    rng = np.random.default_rng(seed)
    table = sample_context_table(kind, n, rng, f32=not f64)
    states = sample_states(kind, n, rng, f32=not f64)
    actions = sample_actions(kind, n, rng)
    ora = OracleClassicEnv(kind, table, applied_mode=(mode == "applied"))
    ora.state[:] = states
    o_ref, r_ref, t_ref, tr_ref, _ = ora.step(actions)
    dt = np.float64 if f64 else np.float32
    st = np.ascontiguousarray(states.astype(dt))
    rows = kernel_rows(kind, table, mode, dt)
    rngs = np.zeros((4, n), dtype=np.uint64)
    sbt = np.zeros(n, dtype=np.uint8)
    el = np.zeros(n, dtype=np.int32)
    o, r, t, tr, _ = hc.step(kind, f64, st, rows, actions, rngs, sbt, el, KINDS[kind]["max_steps"], 0)
    return dict(o=o, r=r, t=t, tr=tr, st=st.astype(np.float64), o_ref=o_ref, r_ref=r_ref, t_ref=t_ref, tr_ref=tr_ref,
                st_ref=ora.state.copy(), table=table)


@pytest.mark.parametrize("kind", KIND_LIST)
def test_single_step_fp32_within_1e5(hc, kind):
    """P1: obs/reward within 1e-5 relative (north star), done masks identical."""
    d = _single_step(hc, kind, f64=False)
    np.testing.assert_allclose(d["o"], d["o_ref"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(d["r"], d["r_ref"], rtol=1e-5, atol=2e-6)
    mism = d["t"] != d["t_ref"]
    if mism.any():  # only legal when the env sits within float32 rounding of a threshold
        assert done_margin(kind, d["st_ref"], d["table"])[mism].max() < 1e-6
    assert mism.sum() <= 1
    assert (d["tr"] == d["tr_ref"]).all()


@pytest.mark.parametrize("kind", KIND_LIST)
def test_single_step_fp64_near_exact(hc, kind):
    d = _single_step(hc, kind, f64=True)
    np.testing.assert_allclose(d["st"], d["st_ref"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(d["o"], d["o_ref"], rtol=2e-7, atol=1e-9)
    assert (d["t"] == d["t_ref"]).all() and (d["tr"] == d["tr_ref"]).all()


def test_cartpole_applied_mode(hc):
    d = _single_step(hc, "cartpole", f64=True, mode="applied")
    np.testing.assert_allclose(d["st"], d["st_ref"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("kind", KIND_LIST)
@pytest.mark.parametrize("f64", [False, True])
def test_reset_bit_exact_stream(hc, kind, f64):
    """P3: device-side reset reproduces the reference's PCG64 stream (gymnasium's discarded draws,
    then CARL's draws) exactly; in fp32 mode the state is the float32 rounding of it."""
    n = 64
    rng = np.random.default_rng(1)
    table = sample_context_table(kind, n, rng, f32=not f64)
    ora = OracleClassicEnv(kind, table)
    o_ref = ora.reset(seed=123)
    dt = np.float64 if f64 else np.float32
    st = np.zeros((n, KINDS[kind]["S"]), dtype=dt)
    rows = kernel_rows(kind, table, "reference", dt)
    rngs = hc.seed_rng(n, 123)
    o = hc.reset(kind, f64, st, rows, rngs)
    np.testing.assert_array_equal(o, o_ref)
    if f64:
        np.testing.assert_array_equal(st, ora.state)
    else:
        np.testing.assert_array_equal(st, ora.state.astype(np.float32))
    # a second reset continues both streams identically
    o_ref2 = ora.reset()
    o2 = hc.reset(kind, f64, st, rows, rngs)
    np.testing.assert_array_equal(o2, o_ref2)


@pytest.mark.parametrize("kind", KIND_LIST)
def test_long_rollout_fp64_with_autoreset(hc, kind):
    """P2 (reference-precision mode): 600 steps incl. terminations, TimeLimit truncations and
    auto-resets; identical done masks, state within 1e-9 (libm vs libm), resets bit-exact."""
    n, T = 48, 600
    rng = np.random.default_rng(2)
    table = sample_context_table(kind, n, rng, f32=False)
    if kind == "acrobot":
        table[: n // 2, 9] = 0.2  # torque noise on half of the envs (env RNG consumed every step)
    max_steps = min(KINDS[kind]["max_steps"], 150)  # several TimeLimit truncations within T steps
    ora = OracleClassicEnv(kind, table, max_steps=max_steps)
    o_ref = ora.reset(seed=9)
    st = np.zeros((n, KINDS[kind]["S"]), dtype=np.float64)
    rows = kernel_rows(kind, table, "reference", np.float64)
    rngs = hc.seed_rng(n, 9)
    o = hc.reset(kind, True, st, rows, rngs)
    np.testing.assert_array_equal(o, o_ref)
    sbt = np.zeros(n, dtype=np.uint8)
    el = np.zeros(n, dtype=np.int32)
    n_done = 0
    for t in range(T):
        a = sample_actions(kind, n, rng)
        o_ref, r_ref, t_ref, tr_ref, fin_ref = ora.step(a, autoreset=True)
        o, r, te, tr, fin = hc.step(kind, True, st, rows, a, rngs, sbt, el, max_steps, 1)
        assert (te == t_ref).all() and (tr == tr_ref).all(), f"done mismatch at step {t}"
        np.testing.assert_allclose(o, o_ref, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(r, r_ref, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(st, ora.state, rtol=1e-9, atol=1e-10)
        done = te | tr
        if done.any():
            np.testing.assert_array_equal(st[done], ora.state[done])  # fresh states are bit-exact
            np.testing.assert_allclose(fin[done], fin_ref[done], rtol=1e-6, atol=1e-7)
            n_done += int(done.sum())
        assert (el == ora.elapsed).all()
    assert n_done > 0


@pytest.mark.parametrize("kind", KIND_LIST)
def test_short_rollout_fp32(hc, kind):
    """P2 (throughput mode): 32 shared-action steps from shared initial states; tolerance scaled
    by the step count (the north star's 1e-5 is a per-step figure; CartPole/Acrobot are chaotic)."""
    n, T = 256, 32
    rng = np.random.default_rng(3)
    table = sample_context_table(kind, n, rng, f32=True)
    ora = OracleClassicEnv(kind, table)
    ora.reset(seed=4)
    ora.state[:] = ora.state.astype(np.float32)
    st = np.ascontiguousarray(ora.state.astype(np.float32))
    rows = kernel_rows(kind, table, "reference", np.float32)
    rngs = hc.seed_rng(n, 4)
    sbt = np.zeros(n, dtype=np.uint8)
    el = np.zeros(n, dtype=np.int32)
    alive = np.ones(n, dtype=bool)
    for t in range(T):
        a = sample_actions(kind, n, rng)
        o_ref, r_ref, t_ref, tr_ref, _ = ora.step(a)
        o, r, te, tr, _ = hc.step(kind, False, st, rows, a, rngs, sbt, el, KINDS[kind]["max_steps"], 0)
        tol = 1e-5 * (t + 1) * (30 if kind == "acrobot" else 4)
        np.testing.assert_allclose(o[alive], o_ref[alive], rtol=tol, atol=tol)
        agree = te == t_ref
        alive &= agree & ~t_ref  # stop comparing an env after its episode ended
    assert alive.sum() > 0


def test_float_threshold_predicates_are_exact(hc):
    """CartPole's done predicate in fp32 mode uses float comparisons that must equal the
    reference's float64 comparison for EVERY float (checked on the neighbourhoods of the thresholds
    and on random floats): done masks stay bit-identical."""
    import ctypes

    rng = np.random.default_rng(0)
    for thr in (2.4, 12 * 2 * np.pi / 360, 0.5, 1.0, 0.1, 3.0000001):
        c = np.float32(thr)
        near = [c]
        for _ in range(50):
            near.append(np.nextafter(near[-1], np.float32(np.inf), dtype=np.float32))
        lo = c
        for _ in range(50):
            lo = np.nextafter(lo, np.float32(-np.inf), dtype=np.float32)
            near.append(lo)
        xs = np.concatenate([np.asarray(near, dtype=np.float32), -np.asarray(near, dtype=np.float32),
                             rng.uniform(-4, 4, 5000).astype(np.float32)])
        n_checked = ctypes.c_int()
        bad = hc.lib.hc_above_below_exact(xs.ctypes.data_as(ctypes.c_void_p), len(xs), ctypes.c_double(thr), ctypes.byref(n_checked))
        assert bad == 0 and n_checked.value == 4 * len(xs)
