"""Parity at BASELINE.json's FULL sizes, every env of the batch against the oracle (VERDICT r01 "weak" #3).

* config 2: CARLCartPole, 65 536 sampled contexts (gravity / length / masscart): seeded reset bit-exact for all 65 536
  PCG64 streams, then teacher-forced steps of the whole batch;
* config 3: CARLPendulum 32 768 + CARLAcrobot 32 768 advanced by ONE mixed launch (`carlb_mixed_step`), both shards
  against their oracles;
* config 4: CARLBraxAnt, 8 192 contexts (gravity / mass_torso / joint_stiffness), teacher-forced env-steps of the whole
  batch against the float32 Brax oracle;
* fp32 done masks on 2^20 random (state, action, context) triples per classic env kind: counted, and every mismatch must
  sit within float32 rounding of its threshold (the float64 mode is bit-identical, tests/test_classic_parity_gpu.py).

Tolerances: obs / reward 1e-5 relative (+2e-6 absolute) in float32 mode -- the north star's; integer streams bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle.classic import FEATURES, KINDS, OracleClassicEnv
from tests.util import done_margin, env_class, sample_actions, sample_context_table, sample_states

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sampled_table(cls, names, n, features, seed=0):
    from carl_b200.context import ContextSampler, UniformFloatContextFeature

    s = ContextSampler([UniformFloatContextFeature(k, lo, hi) for k, (lo, hi) in features.items()], cls.get_context_space(),
                       seed=seed)
    return s.sample_context_table(n, names)


def _teacher_forced_classic(env, ora, kind, steps, rng, label):
    """Re-synchronise the oracle to the device state before every step, step both with the same actions, compare
    EVERY env. Returns the number of fp32 done-mask mismatches (each verified to sit on a threshold)."""
    n = env.num_envs
    mismatches = 0
    for t in range(steps):
        ora.state[:] = env.state.cpu().numpy().astype(np.float64)
        ora.elapsed[:] = env._elapsed.cpu().numpy()
        ora.sbt[:] = env._sbt.cpu().numpy()
        a = sample_actions(kind, n, rng)
        o_ref, r_ref, t_ref, tr_ref, _ = ora.step(a)
        obs, r, te, tr, _ = env.step(torch.from_numpy(a).cuda())
        o = obs["obs"].cpu().numpy() if isinstance(obs, dict) else obs
        te_np, tr_np = te.cpu().numpy(), tr.cpu().numpy()
        done = t_ref | tr_ref | te_np | tr_np  # auto-reset rows carry a fresh state: compared through the reset test
        np.testing.assert_allclose(o[~done], o_ref[~done], rtol=1e-5, atol=2e-6, err_msg=f"{label} step {t} obs")
        np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=1e-5, atol=2e-6, err_msg=f"{label} step {t} reward")
        assert (tr_np == tr_ref).all(), f"{label} step {t}: truncated masks differ"
        mism = te_np != t_ref
        if mism.any():
            assert done_margin(kind, ora.state, ora.ctx)[mism].max() < 1e-6, f"{label} step {t}: done mismatch off-threshold"
            mismatches += int(mism.sum())
    return mismatches


def test_config2_cartpole_65536_reset_and_teacher_forced_steps():
    from carl_b200.envs import CARLCartPole, ContextTable

    n = 65536
    names = FEATURES["cartpole"]
    table = _sampled_table(CARLCartPole, names, n, {"gravity": (5, 15), "length": (0.25, 1.0), "masscart": (0.5, 2.0)})
    table = table.astype(np.float32).astype(np.float64)  # the device holds float32 contexts
    env = CARLCartPole(contexts=ContextTable(names, table), device="cuda:0", autoreset=True)
    ora = OracleClassicEnv("cartpole", table)
    obs, _ = env.reset(seed=0)
    o_ref = ora.reset(seed=0)  # 65 536 numpy Generators: PCG64(SeedSequence(i)), gymnasium's 4 discarded draws, CARL's 4
    np.testing.assert_array_equal(obs["obs"].cpu().numpy(), o_ref)
    np.testing.assert_array_equal(env.state.cpu().numpy(), ora.state.astype(np.float32))
    mism = _teacher_forced_classic(env, ora, "cartpole", 24, np.random.default_rng(1), "config2")
    assert mism == 0, f"{mism} fp32 done-mask mismatches over 24 x 65 536 CartPole steps (the threshold predicates are exact)"


def test_config3_pendulum_plus_acrobot_mixed_launch_vs_oracles():
    from carl_b200.envs import CARLAcrobot, CARLPendulum, ContextTable
    from carl_b200.envs.mixed import MixedBatch

    n = 32768
    tp = _sampled_table(CARLPendulum, FEATURES["pendulum"], n, {"g": (5, 15), "m": (0.5, 2), "l": (0.5, 2)})
    ta = _sampled_table(CARLAcrobot, FEATURES["acrobot"], n, {"LINK_MASS_1": (0.5, 2), "LINK_MASS_2": (0.5, 2),
                                                             "LINK_LENGTH_1": (0.5, 2)}, seed=1)
    tp, ta = tp.astype(np.float32).astype(np.float64), ta.astype(np.float32).astype(np.float64)
    pend = CARLPendulum(contexts=ContextTable(FEATURES["pendulum"], tp), device="cuda:0")
    acro = CARLAcrobot(contexts=ContextTable(FEATURES["acrobot"], ta), device="cuda:0")
    mixed = MixedBatch([pend, acro])
    mixed.reset(seed=3)
    op, oa = OracleClassicEnv("pendulum", tp), OracleClassicEnv("acrobot", ta)
    np.testing.assert_array_equal(pend._obs.cpu().numpy(), op.reset(seed=3))
    np.testing.assert_array_equal(acro._obs.cpu().numpy(), oa.reset(seed=3))
    rng = np.random.default_rng(4)
    acro_mism = 0
    for t in range(12):
        for env, ora in ((pend, op), (acro, oa)):
            ora.state[:] = env.state.cpu().numpy().astype(np.float64)
            ora.elapsed[:] = env._elapsed.cpu().numpy()
        a_p, a_a = sample_actions("pendulum", n, rng), sample_actions("acrobot", n, rng)
        ref_p, ref_a = op.step(a_p), oa.step(a_a)
        out_p, out_a = mixed.step([torch.from_numpy(a_p).cuda(), torch.from_numpy(a_a).cuda()])  # ONE launch
        for (state, r, te, tr, _), (o_ref, r_ref, t_ref, tr_ref, _), kind, ora in ((out_p, ref_p, "pendulum", op),
                                                                                  (out_a, ref_a, "acrobot", oa)):
            np.testing.assert_allclose(state["obs"].cpu().numpy(), o_ref, rtol=1e-5, atol=2e-6, err_msg=f"{kind} step {t}")
            np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=1e-5, atol=2e-6)
            assert (tr.cpu().numpy() == tr_ref).all()
            mism = te.cpu().numpy() != t_ref
            if mism.any():
                assert done_margin(kind, ora.state, ora.ctx)[mism].max() < 1e-6
                acro_mism += int(mism.sum())
    assert acro_mism <= 2


def test_config4_ant_8192_teacher_forced_env_steps():
    """Whole-batch parity of the BASELINE config: gravity ~ U(-15, -5), mass_torso ~ U(5, 20), joint_stiffness scale
    ~ U(0.5, 2) (the v0 feature name BASELINE.json uses; an extension here), context_mode="applied"."""
    from carl_b200.envs import CARLBraxAnt, ContextTable
    from oracle.brax import OracleBraxEnv
    from tests.brax_util import assert_close_scaled, random_q

    n = 8192
    names = list(CARLBraxAnt.get_default_context().keys()) + ["joint_stiffness"]
    d = dict(CARLBraxAnt.get_default_context(), joint_stiffness=1.0)
    rng = np.random.default_rng(0)
    table = np.tile(np.array([float(d[k]) for k in names]), (n, 1))
    table[:, names.index("gravity")] = rng.uniform(-15, -5, n)
    table[:, names.index("mass_torso")] = rng.uniform(5, 20, n)
    table[:, names.index("joint_stiffness")] = rng.uniform(0.5, 2.0, n)
    env = CARLBraxAnt(contexts=ContextTable(names, table), device="cuda:0", context_mode="applied", max_episode_steps=6)
    q, qd = random_q(env._sysd, n, rng)
    env.reset_from_q(q, qd)
    ora = OracleBraxEnv(env._sysd, env._ctx.cpu().numpy(), max_steps=6, autoreset=True)
    ora.init_from_q(q, qd)
    np.testing.assert_allclose(env.state.cpu().numpy(), ora.state, rtol=1e-6, atol=1e-6)
    worst, beyond, total = 0.0, 0, 0
    for t in range(8):  # includes the truncation step (episode length 6): done -> AutoReset to the stored first state
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        ora.state[:] = env.state.cpu().numpy()
        ora.elapsed[:] = env._elapsed.cpu().numpy()
        o_ref, r_ref, d_ref, _ = ora.step(a)
        obs, r, te, tr, _ = env.step(torch.from_numpy(a).cuda())
        assert (te.cpu().numpy() == d_ref).all(), f"done masks differ at step {t}"
        got = obs["obs"].cpu().numpy().astype(np.float64)
        scale = np.maximum(1.0, np.abs(o_ref).max(axis=1, keepdims=True))
        rel = np.abs(got - o_ref) / scale
        worst = max(worst, float(rel.max()))
        beyond += int((rel.max(axis=1) > 1e-5).sum())
        total += n
        np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=1e-4, atol=2e-4)
    # 65 536 env-steps of a contact-rich body: the north-star 1e-5 (relative to the env's own obs magnitude) holds for
    # all but a handful of env-steps in which a contact / joint-limit branch sits within float32 rounding of
    # switching (the impulse is then applied in one arithmetic and not in the other for ONE 5 ms substep); those stay
    # within 1e-4. Both bounds are fixed; the measured numbers go into the log.
    print(f"[config4] Ant 8192 x 8 env-steps: worst obs error {worst:.2e} (relative to the env's obs magnitude), "
          f"{beyond} of {total} env-steps beyond 1e-5")
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "brax_parity_floor.txt"), "a") as f:
            f.write(f"[config4 ant 8192x8 teacher-forced] worst obs err {worst:.2e}; env-steps beyond 1e-5: {beyond}/{total}\n")
    assert beyond <= total // 1000, f"{beyond} of {total} env-steps beyond 1e-5"
    assert worst <= 1e-4, worst


@pytest.mark.parametrize("kind", list(KINDS))
def test_fp32_done_masks_on_a_million_triples(kind):
    """fp32 throughput mode computes the state in float32, the reference in float64: a state that lands within float32
    rounding of a termination threshold can fall on the other side. Count those on 2^20 random triples; every one
    must have a threshold margin < 1e-6. The counts go into gpurun_out/done_mask_counts.json (committed under
    profiles/ and quoted by bench.py)."""
    n = 1 << 20
    rng = np.random.default_rng(2024)
    table = sample_context_table(kind, n, rng)
    states = sample_states(kind, n, rng)
    actions = sample_actions(kind, n, rng)
    from carl_b200.envs import ContextTable

    env = env_class(kind)(contexts=ContextTable(FEATURES[kind], table), device="cuda:0")
    env.reset(seed=0)
    env.state.copy_(torch.from_numpy(states).to(env.state.dtype))
    ora = OracleClassicEnv(kind, table)
    ora.state[:] = states
    _, _, t_ref, _, _ = ora.step(actions)
    _, _, te, _, _ = env.step(torch.from_numpy(actions).cuda())
    mism = te.cpu().numpy() != t_ref
    count = int(mism.sum())
    if count:
        assert done_margin(kind, ora.state, table)[mism].max() < 1e-6
    assert count <= 8, f"{kind}: {count} fp32 done-mask mismatches in 2^20 triples"
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        p = os.path.join(out, "done_mask_counts.json")
        try:
            cur = json.load(open(p))
        except Exception:
            cur = {"triples_per_kind": n, "what": "fp32-mode terminated flags that differ from the float64 oracle on random "
                   "(state, action, context) triples; each within 1e-6 of a termination threshold; float64 mode: 0 by construction",
                   "counts": {}, "terminated_true": {}}
        cur["counts"][kind] = count
        cur["terminated_true"][kind] = int(t_ref.sum())
        json.dump(cur, open(p, "w"))
