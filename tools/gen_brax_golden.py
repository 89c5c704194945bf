"""Dump golden trajectories from a REAL Brax install (brax==0.12.1 + jax), to close the
"parity unpinned" gap of the Brax path. Cannot run in the build container (brax/jax are absent);
run it wherever the reference's own dependencies are installed:

    python tools/gen_brax_golden.py --out tests/golden/brax

For each of ant / halfcheetah / hopper it writes ``<env>.npz`` with: q0, qd0 (the state after
``env.reset``), the action sequence, per-step obs / reward / done, and the spring-backend system
constants (link masses, inertias, joint frames, the <custom> tunables) so that
``tests/test_brax_golden.py`` can (1) overwrite the tunables of ``carl_b200.envs.brax_system`` and
(2) compare the CUDA path step by step (teacher-forced from the dumped states).
"""
import argparse
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/brax")
    ap.add_argument("--steps", type=int, default=64)
    args = ap.parse_args()
    import jax
    import jax.numpy as jp
    from brax import envs

    os.makedirs(args.out, exist_ok=True)
    for name in ("ant", "halfcheetah", "hopper", "walker2d", "inverted_pendulum", "inverted_double_pendulum", "reacher",
                 "humanoid", "humanoidstandup", "pusher"):
        env = envs.create(env_name=name, backend="spring", auto_reset=False, episode_length=1000)
        sys = env.sys
        state = jax.jit(env.reset)(jax.random.PRNGKey(0))
        step = jax.jit(env.step)
        rng = np.random.default_rng(0)
        acts = rng.uniform(-1, 1, (args.steps, env.action_size)).astype(np.float32)
        rec = dict(q0=np.asarray(state.pipeline_state.q), qd0=np.asarray(state.pipeline_state.qd), actions=acts,
                   obs0=np.asarray(state.obs), obs=[], reward=[], done=[], q=[], qd=[], x_pos=[], x_rot=[], xd_vel=[],
                   xd_ang=[])
        for a in acts:
            state = step(state, jp.asarray(a))
            ps = state.pipeline_state
            rec["obs"].append(np.asarray(state.obs)); rec["reward"].append(float(state.reward)); rec["done"].append(float(state.done))
            rec["q"].append(np.asarray(ps.q)); rec["qd"].append(np.asarray(ps.qd))
            rec["x_pos"].append(np.asarray(ps.x.pos)); rec["x_rot"].append(np.asarray(ps.x.rot))
            rec["xd_vel"].append(np.asarray(ps.xd.vel)); rec["xd_ang"].append(np.asarray(ps.xd.ang))
        consts = dict(
            dt=float(sys.opt.timestep), n_frames=int(env._n_frames), gravity=np.asarray(sys.gravity),
            link_mass=np.asarray(sys.link.inertia.mass), link_inertia=np.asarray(sys.link.inertia.i),
            link_com=np.asarray(sys.link.inertia.transform.pos), link_pos=np.asarray(sys.link.transform.pos),
            link_rot=np.asarray(sys.link.transform.rot), joint_pos=np.asarray(sys.link.joint.pos),
            constraint_stiffness=np.asarray(sys.link.constraint_stiffness),
            constraint_vel_damping=np.asarray(sys.link.constraint_vel_damping),
            constraint_limit_stiffness=np.asarray(sys.link.constraint_limit_stiffness),
            constraint_ang_damping=np.asarray(sys.link.constraint_ang_damping),
            baumgarte_erp=float(sys.baumgarte_erp), vel_damping=float(sys.vel_damping), ang_damping=float(sys.ang_damping),
            spring_mass_scale=float(sys.spring_mass_scale), spring_inertia_scale=float(sys.spring_inertia_scale),
            init_q=np.asarray(sys.init_q), gear=np.asarray(sys.actuator.gear),
        )
        out = {k: np.asarray(v) for k, v in rec.items()}
        out.update({f"sys_{k}": v for k, v in consts.items()})
        np.savez(os.path.join(args.out, f"{name}.npz"), **out)
        print("wrote", name)


if __name__ == "__main__":
    main()
