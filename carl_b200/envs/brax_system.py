"""Brax system tables for the spring pipeline (host-side model building).

The reference rebuilds its ``brax.System`` from the MJCF assets inside the brax package on every
context change (``mjcf.load(epath.resource_path("brax") / asset_path)``,
``carl/envs/brax/carl_brax_env.py:271-272``). brax 0.12.1 and its assets are NOT in the reference
tree, so the three in-scope bodies are restated here from the standard Gym/MuJoCo models that
Brax ships: the geometry is *pinned* by CARL's own per-link mass defaults
(``carl_halfcheetah.py:40-57``, ``carl_hopper.py:40-48``), which are the masses MuJoCo derives from
exactly these geoms -- ``tests/test_brax_system.py`` checks every one of them to 7 digits.

What is NOT pinned by anything in the reference (and is therefore exposed as named tunables that
a golden dump from a real Brax install can overwrite, ``tools/gen_brax_golden.py``): the Brax
``<custom>`` constants of the spring backend (constraint stiffness / damping, baumgarte_erp,
spring_mass_scale, spring_inertia_scale, ang_damping).

The packed ``float32`` table built here is uploaded once per handle (``carlb_brax_set_system``)
and staged by every CTA into shared memory with one TMA bulk copy.
"""
from __future__ import annotations

import numpy as np

MAX_LINKS = 12
MAX_POINTS = 32
MAX_Q = 24
HEADER = 32
LINK_STRIDE = 40
POINT_STRIDE = 8
P_SCHED = 7  # point row slot 7: contact schedule (the candidate handled in this slot of the contact passes)
OFF_LINKS = HEADER
OFF_POINTS = OFF_LINKS + LINK_STRIDE * MAX_LINKS
OFF_INIT_Q = OFF_POINTS + POINT_STRIDE * MAX_POINTS
# per-link rows of the 2- and 3-dof revolute joints (humanoid): what dofs 1 and 2 need beyond the link row
DOF_STRIDE = 16
OFF_DOF = OFF_INIT_Q + MAX_Q
# body-vs-body contact pairs (pusher: the gripper's capsules against the pushed ball) and what the pusher's env layer
# reads: a small header (pair count, height of the contact plane, the three links of the observation) + pair rows
MAX_PAIRS = 4
PAIR_STRIDE = 16
PAIR_HEADER = 8
OFF_PAIR = OFF_DOF + DOF_STRIDE * MAX_LINKS
TABLE_FLOATS = OFF_PAIR + PAIR_HEADER + PAIR_STRIDE * MAX_PAIRS  # 1056 floats = 4224 B (a multiple of 16 B for the bulk copy)
(X_N_PAIRS, X_PLANE_Z, X_OBS_LINK0, X_OBS_LINK1, X_OBS_LINK2) = range(5)
# pair row: capsule side (link, candidate row receiving its impulse, segment end points, radius), sphere side (link,
# candidate row, centre, radius)
(R_LINK_A, R_ROW_A, R_A0, R_A1, R_RADIUS_A, R_LINK_B, R_ROW_B, R_B0, R_RADIUS_B) = (0, 1, 2, 5, 8, 9, 10, 11, 14)
# dof-row slots: actuator index / gear / range of dofs 1 and 2, then the sign of each dof's coordinate against the
# right-handed joint frame (x = first axis, y = second axis, z = x cross y; -1 where the MJCF axis is -z)
(D_ACT1, D_ACT2, D_GEAR1, D_GEAR2, D_LO1, D_HI1, D_LO2, D_HI2, D_SIGN0, D_SIGN1, D_SIGN2) = range(11)

# header slots
H_N_LINKS, H_N_Q, H_N_QD, H_N_POINTS, H_N_FRAMES, H_DT, H_ENV, H_N_ACT = range(8)
(H_STIFFNESS, H_VEL_DAMPING_C, H_LIMIT_STIFFNESS, H_ANG_DAMPING_C, H_BAUMGARTE, H_VEL_DAMPING, H_MASS_SCALE,
 H_INERTIA_SCALE) = range(8, 16)
(H_RESET_NOISE, H_CTRL_COST, H_HEALTHY_REWARD, H_HEALTHY_Z_MIN, H_HEALTHY_Z_MAX, H_FORWARD_WEIGHT, H_ANGLE_MIN,
 H_ANGLE_MAX, H_EXCLUDE_POS, H_QD_CLIP, H_TERMINATE, H_MAX_CHILD_POINTS, H_QD_UNIFORM, H_SITE_LINK, H_QD_NOISE,
 H_ACT_SCALE) = range(16, 32)
TUNABLE_NAMES = ["constraint_stiffness", "constraint_vel_damping", "constraint_limit_stiffness",
                 "constraint_ang_damping", "baumgarte_erp", "vel_damping", "spring_mass_scale",
                 "spring_inertia_scale"]

# link slots
(L_PARENT, L_TYPE, L_QIDX, L_QDIDX) = range(4)
L_TPOS, L_TROT, L_JPOS, L_JROT, L_LIM_LO, L_LIM_HI = 4, 7, 11, 14, 18, 19
L_COM, L_IROT, L_IDIAG, L_MASS, L_GEAR, L_ACT, L_CTRL_LO, L_CTRL_HI, L_FIRST_PT, L_N_PT = 20, 23, 27, 30, 31, 32, 33, 34, 35, 36
L_SITE = 37  # body-fixed point read by the env layer (pendulum tip / reacher fingertip), on the H_SITE_LINK row
# TYPE_SLIDE: one prismatic dof along the joint axis; TYPE_SLIDE2: two (joint x and y axes: the reacher's target)
# TYPE_HINGE2 / TYPE_HINGE3: two / three stacked revolute dofs about the joint frame's x, y(, z) axes (humanoid)
TYPE_FREE, TYPE_HINGE, TYPE_SLIDE, TYPE_PLANAR, TYPE_SLIDE2, TYPE_HINGE2, TYPE_HINGE3 = 0, 1, 2, 3, 4, 5, 6
(ENV_ANT, ENV_HALFCHEETAH, ENV_HOPPER, ENV_WALKER2D, ENV_INVERTED_PENDULUM, ENV_INVERTED_DOUBLE_PENDULUM, ENV_REACHER,
 ENV_HUMANOID, ENV_HUMANOIDSTANDUP, ENV_PUSHER) = range(10)
TYPE_DOFS = {TYPE_FREE: (7, 6), TYPE_HINGE: (1, 1), TYPE_SLIDE: (1, 1), TYPE_PLANAR: (3, 3), TYPE_SLIDE2: (2, 2),
             TYPE_HINGE2: (2, 2), TYPE_HINGE3: (3, 3)}
UNLIMITED = 1e30  # joint range of an unlimited hinge


# ----------------------------------------------------------------------------- math
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def quat_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def mat_to_quat(m):
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def frame_with_x(axis):
    """Rotation whose x column is ``axis`` (the joint frame: first dof axis -> x)."""
    x = np.asarray(axis, dtype=np.float64)
    x = x / np.linalg.norm(x)
    helper = np.array([0.0, 0.0, 1.0]) if abs(x[2]) < 0.9 else np.array([0.0, 1.0, 0.0])
    y = np.cross(helper, x)
    y /= np.linalg.norm(y)
    z = np.cross(x, y)
    return mat_to_quat(np.stack([x, y, z], axis=1))


def frame_with_axes(a0, a1):
    """Right-handed joint frame of a 2- / 3-dof revolute joint: x = first axis, y = second axis, z = x cross y
    (the MJCF axes of such a stack are mutually orthogonal; a third axis is +-z)."""
    x = np.asarray(a0, dtype=np.float64)
    y = np.asarray(a1, dtype=np.float64)
    x, y = x / np.linalg.norm(x), y / np.linalg.norm(y)
    assert abs(np.dot(x, y)) < 1e-12, "stacked hinge axes must be orthogonal"
    return mat_to_quat(np.stack([x, y, np.cross(x, y)], axis=1))


# ------------------------------------------------------------------------- geometry
def capsule(p0, p1, r):
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    return dict(type="capsule", p0=p0, p1=p1, r=float(r))


def capsule_axis_angle(pos, axis, angle, r, half):
    """MJCF capsule given by pos + axisangle + size (radius, half-length); local axis is z."""
    d = quat_to_mat(quat_axis_angle(axis, angle)) @ np.array([0.0, 0.0, half])
    pos = np.asarray(pos, float)
    return capsule(pos - d, pos + d, r)


def sphere(pos, r):
    return dict(type="sphere", p0=np.asarray(pos, float), r=float(r))


def geom_mass_inertia(g, density):
    """MuJoCo ``inertiafromgeom``: mass, COM and inertia tensor (about the COM, link frame)."""
    r = g["r"]
    if g["type"] == "sphere":
        m = density * 4.0 / 3.0 * np.pi * r**3
        return m, g["p0"], np.eye(3) * (0.4 * m * r * r)
    axis = g["p1"] - g["p0"]
    h = np.linalg.norm(axis)
    z = axis / h
    m_c = density * np.pi * r * r * h
    m_s = density * 4.0 / 3.0 * np.pi * r**3
    i_ax = m_c * r * r / 2 + m_s * 0.4 * r * r
    i_lat = m_c * (r * r / 4 + h * h / 12) + m_s * (0.4 * r * r + h * h / 4 + 3 * h * r / 8)
    zz = np.outer(z, z)
    inertia = i_lat * (np.eye(3) - zz) + i_ax * zz
    return m_c + m_s, 0.5 * (g["p0"] + g["p1"]), inertia


def body_inertia(geoms, density):
    parts = [geom_mass_inertia(g, g.get("density", density)) for g in geoms]
    m = sum(p[0] for p in parts)
    com = sum(p[0] * p[1] for p in parts) / m
    inertia = np.zeros((3, 3))
    for mi, ci, ii in parts:
        d = ci - com
        inertia += ii + mi * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    w, v = np.linalg.eigh(inertia)
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
    if np.allclose(inertia, np.diag(np.diag(inertia)), atol=1e-12 * max(1.0, np.abs(inertia).max())):
        w, v = np.diag(inertia).copy(), np.eye(3)  # already principal: keep the link axes
    return m, com, mat_to_quat(v), w


def contact_points(geoms):
    pts = []
    for g in geoms:
        if not g.get("collide", True):  # contype = conaffinity = 0
            continue
        if g["type"] == "sphere":
            pts.append((g["p0"], g["r"], g.get("friction")))
        else:  # capsule vs plane: both end spheres are contact candidates
            pts.append((g["p0"], g["r"], g.get("friction")))
            pts.append((g["p1"], g["r"], g.get("friction")))
    return pts


# --------------------------------------------------------------------------- models
def _link(name, parent, typ, pos, geoms, axis=None, joint_pos=(0, 0, 0), limit=(0, 0), gear=0.0, quat=(1, 0, 0, 0),
          ctrl_range=(-1.0, 1.0)):
    """One link and the joint to its parent. For TYPE_HINGE2 / TYPE_HINGE3 ``axis``, ``limit`` and ``gear`` are
    lists with one entry per dof (MJCF order of the stacked <joint> elements)."""
    quat = np.asarray(quat, float)
    return dict(name=name, parent=parent, type=typ, pos=np.asarray(pos, float), quat=quat / np.linalg.norm(quat),
                geoms=geoms, axis=axis, joint_pos=np.asarray(joint_pos, float), limit=limit, gear=gear, ctrl_range=ctrl_range)


def ant_model():
    """Brax/Gym ``ant.xml``: torso sphere + 4 x (aux capsule fused into the torso link, hip link,
    ankle link); hip hinge about z (+-30 deg), ankle hinge about (-+1, 1, 0) (30..70 deg, mirrored)."""
    r = 0.08
    deg = np.pi / 180
    torso_geoms = [sphere((0, 0, 0), 0.25)] + [capsule((0, 0, 0), (sx * 0.2, sy * 0.2, 0), r)
                                              for sx, sy in ((1, 1), (-1, 1), (-1, -1), (1, -1))]
    links = [_link("torso", -1, TYPE_FREE, (0, 0, 0.75), torso_geoms)]
    legs = [  # (sx, sy, ankle axis, ankle range)
        (1, 1, (-1, 1, 0), (30, 70)), (-1, 1, (1, 1, 0), (-70, -30)),
        (-1, -1, (-1, 1, 0), (-70, -30)), (1, -1, (1, 1, 0), (30, 70))]
    for i, (sx, sy, aax, arange) in enumerate(legs, start=1):
        hip = len(links)
        links.append(_link(f"aux_{i}", 0, TYPE_HINGE, (sx * 0.2, sy * 0.2, 0), [capsule((0, 0, 0), (sx * 0.2, sy * 0.2, 0), r)],
                           axis=(0, 0, 1), limit=(-30 * deg, 30 * deg), gear=150.0))
        links.append(_link(f"ankle_{i}", hip, TYPE_HINGE, (sx * 0.2, sy * 0.2, 0), [capsule((0, 0, 0), (sx * 0.4, sy * 0.4, 0), r)],
                           axis=aax, limit=(arange[0] * deg, arange[1] * deg), gear=150.0))
    init_q = np.array([0, 0, 0.55, 1, 0, 0, 0, 0, 1, 0, -1, 0, -1, 0, 1], dtype=np.float64)
    return dict(
        name="ant", env=ENV_ANT, links=links, density=5.0, total_mass=None, friction=1.0, init_q=init_q,
        dt=0.005, n_frames=10,
        tunables=dict(constraint_stiffness=4000.0, constraint_vel_damping=20.0, constraint_limit_stiffness=1000.0,
                      constraint_ang_damping=10.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=1.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.1, ctrl_cost=0.5, healthy_reward=1.0, z_min=0.2, z_max=1.0, forward_weight=1.0,
                        angle_min=0.0, angle_max=0.0, exclude_pos=2, qd_clip=0.0, terminate=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0,
        # actuator order of the MJCF <actuator> block: hip_4, ankle_4, hip_1, ankle_1, hip_2, ankle_2, hip_3, ankle_3
        actuator_links=["aux_4", "ankle_4", "aux_1", "ankle_1", "aux_2", "ankle_2", "aux_3", "ankle_3"],
    )


def halfcheetah_model():
    """Gym ``half_cheetah.xml`` (angles in radian, ``settotalmass=14``, friction 0.4)."""
    r = 0.046
    y = (0, 1, 0)
    links = [
        _link("torso", -1, TYPE_PLANAR, (0, 0, 0.7),
              [capsule((-0.5, 0, 0), (0.5, 0, 0), r), capsule_axis_angle((0.6, 0, 0.1), y, 0.87, r, 0.15)], axis=y),
        _link("bthigh", 0, TYPE_HINGE, (-0.5, 0, 0), [capsule_axis_angle((0.1, 0, -0.13), y, -3.8, r, 0.145)], axis=y,
              limit=(-0.52, 1.05), gear=120.0),
        _link("bshin", 1, TYPE_HINGE, (0.16, 0, -0.25), [capsule_axis_angle((-0.14, 0, -0.07), y, -2.03, r, 0.15)], axis=y,
              limit=(-0.785, 0.785), gear=90.0),
        _link("bfoot", 2, TYPE_HINGE, (-0.28, 0, -0.14), [capsule_axis_angle((0.03, 0, -0.097), y, -0.27, r, 0.094)], axis=y,
              limit=(-0.4, 0.785), gear=60.0),
        _link("fthigh", 0, TYPE_HINGE, (0.5, 0, 0), [capsule_axis_angle((-0.07, 0, -0.12), y, 0.52, r, 0.133)], axis=y,
              limit=(-1.0, 0.7), gear=120.0),
        _link("fshin", 4, TYPE_HINGE, (-0.14, 0, -0.24), [capsule_axis_angle((0.065, 0, -0.09), y, -0.6, r, 0.106)], axis=y,
              limit=(-1.2, 0.87), gear=100.0),  # Brax spring/positional gear override [120,90,60,120,100,100]
        _link("ffoot", 5, TYPE_HINGE, (0.13, 0, -0.18), [capsule_axis_angle((0.045, 0, -0.07), y, -0.6, r, 0.07)], axis=y,
              limit=(-0.5, 0.5), gear=100.0),
    ]
    return dict(
        name="halfcheetah", env=ENV_HALFCHEETAH, links=links, density=1000.0, total_mass=14.0, friction=0.4,
        init_q=np.zeros(9), dt=0.003125, n_frames=16,
        tunables=dict(constraint_stiffness=25000.0, constraint_vel_damping=80.0, constraint_limit_stiffness=2000.0,
                      constraint_ang_damping=10.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.1, ctrl_cost=0.1, healthy_reward=0.0, z_min=-1e9, z_max=1e9, forward_weight=1.0,
                        angle_min=0.0, angle_max=0.0, exclude_pos=1, qd_clip=0.0, terminate=0.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0,
        actuator_links=["bthigh", "bshin", "bfoot", "fthigh", "fshin", "ffoot"],
    )


def hopper_model():
    """Gym ``hopper.xml`` in local coordinates (density 1000; joints about (0,-1,0); gear 200)."""
    ax = (0, -1, 0)
    deg = np.pi / 180
    foot = capsule((-0.13, 0, 0), (0.26, 0, 0), 0.06)
    foot["friction"] = 2.0
    links = [
        _link("torso", -1, TYPE_PLANAR, (0, 0, 1.25), [capsule((0, 0, 0.2), (0, 0, -0.2), 0.05)], axis=(0, 1, 0)),
        _link("thigh", 0, TYPE_HINGE, (0, 0, -0.2), [capsule((0, 0, 0), (0, 0, -0.45), 0.05)], axis=ax,
              limit=(-150 * deg, 0.0), gear=200.0),
        _link("leg", 1, TYPE_HINGE, (0, 0, -0.45), [capsule((0, 0, 0), (0, 0, -0.5), 0.04)], axis=ax,
              limit=(-150 * deg, 0.0), gear=200.0),
        _link("foot", 2, TYPE_HINGE, (0, 0, -0.5), [foot], axis=ax, limit=(-45 * deg, 45 * deg), gear=200.0),
    ]
    return dict(
        name="hopper", env=ENV_HOPPER, links=links, density=1000.0, total_mass=None, friction=1.0,
        init_q=np.zeros(6), dt=0.002, n_frames=4,
        tunables=dict(constraint_stiffness=30000.0, constraint_vel_damping=100.0, constraint_limit_stiffness=2000.0,
                      constraint_ang_damping=10.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=5e-3, ctrl_cost=1e-3, healthy_reward=1.0, z_min=0.7, z_max=1e9, forward_weight=1.0,
                        angle_min=-0.2, angle_max=0.2, exclude_pos=1, qd_clip=10.0, terminate=1.0, qd_uniform=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0,
        actuator_links=["thigh", "leg", "foot"],
    )


def walker2d_model():
    """Gym ``walker2d.xml`` in local coordinates (density 1000; hinges about (0,-1,0); gear 100).
    The geometry is pinned by CARL's mass defaults (``carl/envs/brax/carl_walker2d.py:37-57``)."""
    ax = (0, -1, 0)
    deg = np.pi / 180

    def leg(suffix, parent_of_thigh, foot_friction):
        foot = capsule((0, 0, 0), (0.2, 0, 0), 0.06)
        foot["friction"] = foot_friction
        base = 1 if suffix == "" else 4
        return [
            _link(f"thigh{suffix}", parent_of_thigh, TYPE_HINGE, (0, 0, -0.2), [capsule((0, 0, 0), (0, 0, -0.45), 0.05)], axis=ax,
                  limit=(-150 * deg, 0.0), gear=100.0),
            _link(f"leg{suffix}", base, TYPE_HINGE, (0, 0, -0.45), [capsule((0, 0, 0), (0, 0, -0.5), 0.04)], axis=ax,
                  limit=(-150 * deg, 0.0), gear=100.0),
            _link(f"foot{suffix}", base + 1, TYPE_HINGE, (0, 0, -0.5), [foot], axis=ax, limit=(-45 * deg, 45 * deg), gear=100.0),
        ]

    links = [_link("torso", -1, TYPE_PLANAR, (0, 0, 1.25), [capsule((0, 0, 0.2), (0, 0, -0.2), 0.05)], axis=(0, 1, 0))]
    links += leg("", 0, 0.9) + leg("_left", 0, 1.9)
    return dict(
        name="walker2d", env=ENV_WALKER2D, links=links, density=1000.0, total_mass=None, friction=1.0,
        init_q=np.zeros(9), dt=0.002, n_frames=4,
        tunables=dict(constraint_stiffness=30000.0, constraint_vel_damping=100.0, constraint_limit_stiffness=2000.0,
                      constraint_ang_damping=10.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=5e-3, ctrl_cost=1e-3, healthy_reward=1.0, z_min=0.8, z_max=2.0, forward_weight=1.0,
                        angle_min=-1.0, angle_max=1.0, exclude_pos=1, qd_clip=10.0, terminate=1.0, qd_uniform=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0,
        actuator_links=["thigh", "leg", "foot", "thigh_left", "leg_left", "foot_left"],
    )


def inverted_pendulum_model():
    """Gym / Brax ``inverted_pendulum.xml``: a cart on a slide joint along x (range +-1) carrying a pole on a
    hinge about y (+-90 deg); motor on the slider, gear 100, ctrlrange +-3; no collision geometry
    (contype = conaffinity = 0). CARL's ``mass_cart`` / ``mass_pole`` defaults of 1 are placeholders, not the
    MJCF-derived masses (``carl/envs/brax/carl_inverted_pendulum.py:27-32``)."""
    deg = np.pi / 180
    links = [
        _link("cart", -1, TYPE_SLIDE, (0, 0, 0), [capsule((-0.1, 0, 0), (0.1, 0, 0), 0.1)], axis=(1, 0, 0), limit=(-1.0, 1.0),
              gear=100.0, ctrl_range=(-3.0, 3.0)),
        _link("pole", 0, TYPE_HINGE, (0, 0, 0), [capsule((0, 0, 0), (0.001, 0, 0.6), 0.049)], axis=(0, 1, 0),
              limit=(-90 * deg, 90 * deg)),
    ]
    return dict(
        name="inverted_pendulum", env=ENV_INVERTED_PENDULUM, links=links, density=1000.0, total_mass=None, friction=1.0,
        init_q=np.zeros(2), dt=0.005, n_frames=4, contacts=False,
        tunables=dict(constraint_stiffness=10000.0, constraint_vel_damping=50.0, constraint_limit_stiffness=10000.0,
                      constraint_ang_damping=0.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.01, qd_noise=0.01, ctrl_cost=0.0, healthy_reward=1.0, z_min=-1e9, z_max=1e9,
                        forward_weight=0.0, angle_min=-0.2, angle_max=0.2, exclude_pos=0, qd_clip=0.0, terminate=1.0,
                        qd_uniform=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0, actuator_links=["cart"],
    )


def inverted_double_pendulum_model():
    """Gym / Brax ``inverted_double_pendulum.xml``: cart (slide x, +-1) + two 0.6 m poles on unlimited hinges about y;
    motor on the slider, gear 500, ctrlrange +-1; tip site at (0, 0, 0.6) of the second pole; no collisions."""
    links = [
        _link("cart", -1, TYPE_SLIDE, (0, 0, 0), [capsule((-0.1, 0, 0), (0.1, 0, 0), 0.1)], axis=(1, 0, 0), limit=(-1.0, 1.0),
              gear=500.0),
        _link("pole", 0, TYPE_HINGE, (0, 0, 0), [capsule((0, 0, 0), (0, 0, 0.6), 0.045)], axis=(0, 1, 0),
              limit=(-UNLIMITED, UNLIMITED)),
        _link("pole2", 1, TYPE_HINGE, (0, 0, 0.6), [capsule((0, 0, 0), (0, 0, 0.6), 0.045)], axis=(0, 1, 0),
              limit=(-UNLIMITED, UNLIMITED)),
    ]
    return dict(
        name="inverted_double_pendulum", env=ENV_INVERTED_DOUBLE_PENDULUM, links=links, density=1000.0, total_mass=None,
        friction=1.0, init_q=np.zeros(3), dt=0.005, n_frames=4, contacts=False, site=("pole2", (0.0, 0.0, 0.6)), obs_dim=8,
        tunables=dict(constraint_stiffness=10000.0, constraint_vel_damping=50.0, constraint_limit_stiffness=10000.0,
                      constraint_ang_damping=0.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.01, qd_noise=0.1, ctrl_cost=0.0, healthy_reward=10.0, z_min=-1e9, z_max=1e9,
                        forward_weight=0.0, angle_min=0.0, angle_max=0.0, exclude_pos=0, qd_clip=10.0, terminate=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0, actuator_links=["cart"],
    )


def reacher_model():
    """Gym / Brax ``reacher.xml``: two 0.1 m capsule links on hinges about z (joint0 unlimited, joint1 +-3 rad), the
    fingertip sphere fused into the second link, and a target sphere on two slide joints (x, y). Gears 25 (Brax's
    spring / positional override of the MJCF's 200). The capsule geometry is pinned by CARL's ``mass_body0`` /
    ``mass_body1`` defaults (``carl/envs/brax/carl_reacher.py:36-41``)."""
    z = (0, 0, 1)
    links = [
        _link("body0", -1, TYPE_HINGE, (0, 0, 0.01), [capsule((0, 0, 0), (0.1, 0, 0), 0.01)], axis=z,
              limit=(-UNLIMITED, UNLIMITED), gear=25.0),
        _link("body1", 0, TYPE_HINGE, (0.1, 0, 0), [capsule((0, 0, 0), (0.1, 0, 0), 0.01), sphere((0.11, 0, 0), 0.01)], axis=z,
              limit=(-3.0, 3.0), gear=25.0),
        _link("target", -1, TYPE_SLIDE2, (0, 0, 0.01), [sphere((0, 0, 0), 0.009)], axis=(1, 0, 0), limit=(-0.27, 0.27)),
    ]
    return dict(
        name="reacher", env=ENV_REACHER, links=links, density=1000.0, total_mass=None, friction=1.0,
        init_q=np.zeros(4), dt=0.005, n_frames=4, contacts=False, site=("body1", (0.11, 0.0, 0.0)), obs_dim=11,
        # unit effective masses / inertias (scale 1): the 36 g links would need a far smaller step otherwise
        tunables=dict(constraint_stiffness=5000.0, constraint_vel_damping=50.0, constraint_limit_stiffness=1000.0,
                      constraint_ang_damping=5.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=1.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.1, qd_noise=0.005, ctrl_cost=1.0, healthy_reward=0.0, z_min=-1e9, z_max=1e9,
                        forward_weight=0.0, angle_min=0.0, angle_max=0.0, exclude_pos=0, qd_clip=0.0, terminate=0.0,
                        qd_uniform=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0, actuator_links=["body0", "body1"],
    )


def humanoid_geometry():
    """The collision geometry of Gym / Brax ``humanoid.xml`` per link, pinned by CARL's own mass defaults
    (``carl/envs/brax/carl_humanoid.py:37-75``): ``tests/test_brax_system.py::test_humanoid_geometry_matches_carl_masses``
    reproduces all ten to 7 digits."""
    arm_up = lambda sy: [capsule((0, 0, 0), (0.16, sy * 0.16, -0.16), 0.04)]
    arm_lo = lambda sy: [capsule((0.01, sy * 0.01, 0.01), (0.17, sy * 0.17, 0.17), 0.031), sphere((0.18, sy * 0.18, 0.18), 0.04)]
    thigh = lambda sy: [capsule((0, 0, 0), (0, sy * 0.01, -0.34), 0.06)]
    shin = [capsule((0, 0, 0), (0, 0, -0.3), 0.049), sphere((0, 0, -0.35), 0.075)]
    return {
        "torso": [capsule((0, -0.07, 0), (0, 0.07, 0), 0.07), sphere((0, 0, 0.19), 0.09),
                  capsule((-0.01, -0.06, -0.12), (-0.01, 0.06, -0.12), 0.06)],
        "lwaist": [capsule((0, -0.06, 0), (0, 0.06, 0), 0.06)],
        "pelvis": [capsule((-0.02, -0.07, 0), (-0.02, 0.07, 0), 0.09)],
        "right_thigh": thigh(1), "right_shin": shin, "left_thigh": thigh(-1), "left_shin": shin,
        "right_upper_arm": arm_up(-1), "right_lower_arm": arm_lo(1), "left_upper_arm": arm_up(1), "left_lower_arm": arm_lo(-1),
    }


def humanoid_model():
    """Gym / Brax ``humanoid.xml`` (angles in degree, density 1000): torso on a free joint; abdomen (z, y) and
    abdomen x; hips (x, z, y) and knees; shoulders (two oblique axes) and elbows -- 11 links, 17 actuated dofs,
    ctrlrange +-0.4. Spring backend of brax 0.12.1 ``envs/humanoid.py``: dt 0.0015, 10 substeps, gears 350 (abdomen,
    legs) / 100 (arms). The feet and hands are spheres fused into the shins / lower arms. Every geom is a
    ground-contact candidate (29 end spheres), no body-vs-body contacts. The geometry is pinned by CARL's mass
    defaults (``carl/envs/brax/carl_humanoid.py:37-75``, ``tests/test_brax_system.py``)."""
    deg = np.pi / 180
    g = humanoid_geometry()
    x, y, z = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    rng = lambda lo, hi: (lo * deg, hi * deg)
    leg, arm = 350.0, 100.0
    links = [
        _link("torso", -1, TYPE_FREE, (0, 0, 1.4), g["torso"]),
        _link("lwaist", 0, TYPE_HINGE2, (-0.01, 0, -0.26), g["lwaist"], axis=[z, y], joint_pos=(0, 0, 0.065),
              limit=[rng(-45, 45), rng(-75, 30)], gear=[leg, leg], quat=(1, 0, -0.002, 0)),
        _link("pelvis", 1, TYPE_HINGE, (0, 0, -0.165), g["pelvis"], axis=x, joint_pos=(0, 0, 0.1), limit=rng(-35, 35), gear=leg,
              quat=(1, 0, -0.002, 0)),
        _link("right_thigh", 2, TYPE_HINGE3, (0, -0.1, -0.04), g["right_thigh"], axis=[x, z, y],
              limit=[rng(-25, 5), rng(-60, 35), rng(-110, 20)], gear=[leg] * 3),
        _link("right_shin", 3, TYPE_HINGE, (0, 0.01, -0.403), g["right_shin"], axis=(0, -1, 0), joint_pos=(0, 0, 0.02),
              limit=rng(-160, -2), gear=leg),
        _link("left_thigh", 2, TYPE_HINGE3, (0, 0.1, -0.04), g["left_thigh"], axis=[(-1, 0, 0), (0, 0, -1), y],
              limit=[rng(-25, 5), rng(-60, 35), rng(-120, 20)], gear=[leg] * 3),
        _link("left_shin", 5, TYPE_HINGE, (0, -0.01, -0.403), g["left_shin"], axis=(0, -1, 0), joint_pos=(0, 0, 0.02),
              limit=rng(-160, -2), gear=leg),
        _link("right_upper_arm", 0, TYPE_HINGE2, (0, -0.17, 0.06), g["right_upper_arm"], axis=[(2, 1, 1), (0, -1, 1)],
              limit=[rng(-85, 60), rng(-85, 60)], gear=[arm, arm]),
        _link("right_lower_arm", 7, TYPE_HINGE, (0.18, -0.18, -0.18), g["right_lower_arm"], axis=(0, -1, 1), limit=rng(-90, 50),
              gear=arm),
        _link("left_upper_arm", 0, TYPE_HINGE2, (0, 0.17, 0.06), g["left_upper_arm"], axis=[(2, -1, 1), (0, 1, 1)],
              limit=[rng(-60, 85), rng(-60, 85)], gear=[arm, arm]),
        _link("left_lower_arm", 9, TYPE_HINGE, (0.18, 0.18, -0.18), g["left_lower_arm"], axis=(0, -1, -1), limit=rng(-90, 50),
              gear=arm),
    ]
    for l in links[1:]:
        l["ctrl_range"] = (-0.4, 0.4)
    init_q = np.zeros(24)
    init_q[:7] = (0, 0, 1.4, 1, 0, 0, 0)
    return dict(
        name="humanoid", env=ENV_HUMANOID, links=links, density=1000.0, total_mass=None, friction=1.0, init_q=init_q,
        dt=0.0015, n_frames=10, obs_dim=244,
        tunables=dict(constraint_stiffness=27000.0, constraint_vel_damping=80.0, constraint_limit_stiffness=2500.0,
                      constraint_ang_damping=30.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.01, qd_noise=0.01, ctrl_cost=0.1, healthy_reward=5.0, z_min=1.0, z_max=2.0,
                        forward_weight=1.25, angle_min=0.0, angle_max=0.0, exclude_pos=2, qd_clip=0.0, terminate=1.0,
                        qd_uniform=1.0),
        stock_gravity=-9.81, stock_ang_damping=-0.05, stock_elasticity=0.0,
        # MJCF <actuator> order; "link:k" is dof k of a stacked joint
        actuator_links=["lwaist:1", "lwaist:0", "pelvis", "right_thigh:0", "right_thigh:1", "right_thigh:2", "right_shin",
                        "left_thigh:0", "left_thigh:1", "left_thigh:2", "left_shin", "right_upper_arm:0", "right_upper_arm:1",
                        "right_lower_arm", "left_upper_arm:0", "left_upper_arm:1", "left_lower_arm"],
    )


def _rotate_model_frames(links, quat):
    """Express every link-local vector of a model in link frames turned by ``quat`` (the body keeps its shape;
    the pose that the identity root orientation describes changes)."""
    r = quat_to_mat(np.asarray(quat, float))
    rv = lambda v: r @ np.asarray(v, float)
    out = []
    for l in links:
        l = dict(l)
        l["pos"], l["joint_pos"] = rv(l["pos"]), rv(l["joint_pos"])
        q = np.asarray(l["quat"], float)
        l["quat"] = np.concatenate([[q[0]], rv(q[1:])])
        if l["axis"] is not None:
            l["axis"] = [tuple(rv(a)) for a in l["axis"]] if isinstance(l["axis"], list) else tuple(rv(l["axis"]))
        geoms = []
        for gm in l["geoms"]:
            gm = dict(gm)
            gm["p0"] = rv(gm["p0"])
            if "p1" in gm:
                gm["p1"] = rv(gm["p1"])
            geoms.append(gm)
        l["geoms"] = geoms
        out.append(l)
    return out


def humanoidstandup_model():
    """Gym / Brax ``humanoidstandup.xml``: the humanoid lying on its back, root at z = 0.105 with the identity
    orientation (the XML defines the body along x: head towards -x, legs towards +x). Restated as the humanoid's
    links expressed in frames turned by -90 deg about y; CARL's mass defaults are the humanoid's
    (``carl/envs/brax/carl_humanoidstandup.py:32-69``). Env layer of brax 0.12.1 ``envs/humanoidstandup.py``:
    reward = z_torso / dt + 1 - 0.01 |a|^2, never done."""
    m = humanoid_model()
    m["links"] = _rotate_model_frames(m["links"], quat_axis_angle((0, 1, 0), -np.pi / 2))
    m["links"][0]["pos"] = np.array([0.0, 0.0, 0.105])
    m["name"], m["env"] = "humanoidstandup", ENV_HUMANOIDSTANDUP
    m["init_q"] = m["init_q"].copy()
    m["init_q"][2] = 0.105
    m["env_params"] = dict(m["env_params"], ctrl_cost=0.01, healthy_reward=1.0, z_min=-1e9, z_max=1e9, forward_weight=0.0,
                           terminate=0.0)
    return m


def pusher_model():
    """Gym / Brax ``pusher.xml`` (``arm3d``, density 300, zero gravity): a 7-dof arm (seven single hinges; the
    jointless ``r_upper_arm_link`` / ``r_forearm_link`` / ``tips_arm`` bodies are fused into their parents), the pushed
    object and the goal marker on two slide joints each (MJCF order: y, then x). Collisions as the MJCF's contype /
    conaffinity bits allow: the gripper's three capsules and the object against the table plane, and the three capsules
    against the object. The object is a ball of radius 0.05 whose mass is CARL's ``mass_object`` default (1.8326e-3 kg =
    density 3.5: the Brax asset replaces Gym's cylinder, which the spring backend cannot collide); the arm geometry is
    pinned by CARL's seven link-mass defaults (``carl/envs/brax/carl_pusher.py:36-84``, ``tests/test_brax_system.py``).
    Spring backend of brax 0.12.1 ``envs/pusher.py``: dt 0.001, 50 substeps, gears 20. The world is shifted up by 0.325 so
    that the table is the engine's contact plane z = 0; the observation subtracts the shift again (``X_PLANE_Z``)."""
    zt = 0.325  # the table plane sits at z = -0.325 in the MJCF
    x, y, z = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    no = dict(collide=False)
    nc = lambda g: dict(g, **no)
    grip = [capsule((0, -0.1, 0), (0, 0.1, 0), 0.02), capsule((0, -0.1, 0), (0.1, -0.1, 0), 0.02), capsule((0, 0.1, 0), (0.1, 0.1, 0), 0.02)]
    ball = sphere((0, 0, 0), 0.05)
    ball["density"] = 3.5
    marker = nc(sphere((0, 0, 0), 0.001))   # the goal marker: no collisions, negligible mass
    marker["density"] = 1e-5
    links = [
        _link("r_shoulder_pan_link", -1, TYPE_HINGE, (0, -0.6, zt),
              [nc(sphere((-0.06, 0.05, 0.2), 0.05)), nc(sphere((0.06, 0.05, 0.2), 0.05)), nc(sphere((-0.06, 0.09, 0.2), 0.03)),
               nc(sphere((0.06, 0.09, 0.2), 0.03)), nc(capsule((0, 0, -0.4), (0, 0, 0.2), 0.1))], axis=z, limit=(-2.2854, 1.714602)),
        _link("r_shoulder_lift_link", 0, TYPE_HINGE, (0.1, 0, 0), [nc(capsule((0, -0.1, 0), (0, 0.1, 0), 0.1))], axis=y,
              limit=(-0.5236, 1.3963)),
        _link("r_upper_arm_roll_link", 1, TYPE_HINGE, (0, 0, 0),
              [nc(capsule((-0.1, 0, 0), (0.1, 0, 0), 0.02)), nc(capsule((0, 0, 0), (0.4, 0, 0), 0.06))], axis=x, limit=(-1.5, 1.7)),
        _link("r_elbow_flex_link", 2, TYPE_HINGE, (0.4, 0, 0), [nc(capsule((0, -0.02, 0), (0, 0.02, 0), 0.06))], axis=y,
              limit=(-2.3213, 0.0)),
        _link("r_forearm_roll_link", 3, TYPE_HINGE, (0, 0, 0),
              [nc(capsule((-0.1, 0, 0), (0.1, 0, 0), 0.02)), nc(capsule((0, 0, 0), (0.291, 0, 0), 0.05))], axis=x, limit=(-1.5, 1.5)),
        _link("r_wrist_flex_link", 4, TYPE_HINGE, (0.321, 0, 0), [nc(capsule((0, -0.02, 0), (0, 0.02, 0), 0.01))], axis=y,
              limit=(-1.094, 0.0)),
        _link("r_wrist_roll_link", 5, TYPE_HINGE, (0, 0, 0),
              [nc(sphere((0.1, -0.1, 0), 0.01)), nc(sphere((0.1, 0.1, 0), 0.01))] + grip, axis=x, limit=(-1.5, 1.5)),
        _link("object", -1, TYPE_SLIDE2, (0.45, -0.05, -0.275 + zt), [ball], axis=y, limit=(-10.3213, 10.3)),
        _link("goal", -1, TYPE_SLIDE2, (0.45, -0.05, -0.323 + zt), [marker], axis=y, limit=(-10.3213, 10.3)),
    ]
    for l in links[:7]:
        l["gear"], l["ctrl_range"] = 20.0, (-2.0, 2.0)
    links[7]["axis2"] = links[8]["axis2"] = x  # slide order of the MJCF: y, then x
    return dict(
        name="pusher", env=ENV_PUSHER, links=links, density=300.0, total_mass=None, friction=0.8, init_q=np.zeros(11),
        dt=0.001, n_frames=50, obs_dim=23, plane_z=zt, obs_links=("r_wrist_flex_link", "object", "goal"),
        pairs=[("r_wrist_roll_link", k, "object", 0) for k in (2, 3, 4)],  # (capsule link, geom index, sphere link, geom index)
        # unit effective masses / inertias (scale 1), like the Ant and the reacher: the 2-gram ball and 5-gram wrist
        # link would need a far smaller step otherwise
        tunables=dict(constraint_stiffness=5000.0, constraint_vel_damping=50.0, constraint_limit_stiffness=1000.0,
                      constraint_ang_damping=5.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=1.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.0, qd_noise=0.005, ctrl_cost=0.1, healthy_reward=0.0, z_min=-1e9, z_max=1e9,
                        forward_weight=0.0, angle_min=0.0, angle_max=0.0, exclude_pos=0, qd_clip=0.0, terminate=0.0,
                        qd_uniform=1.0),
        stock_gravity=0.0, stock_ang_damping=-0.05, stock_elasticity=0.0,
        actuator_links=["r_shoulder_pan_link", "r_shoulder_lift_link", "r_upper_arm_roll_link", "r_elbow_flex_link",
                        "r_forearm_roll_link", "r_wrist_flex_link", "r_wrist_roll_link"],
    )


def _initial_point_clearance(links, init_q, pts_all, q_idx):
    """Height above the ground (sphere centre z - radius) of every contact candidate in the initial pose:
    a plain forward-kinematics pass (joints at the link origin, as in every model here). Only used to ORDER
    the candidates for the kernel's contact passes, never for physics."""
    world = []
    for i, l in enumerate(links):
        q = init_q[q_idx[i]:]
        if l["type"] == TYPE_FREE:
            pos, rot = np.asarray(q[:3], float), np.asarray(q[3:7], float) / np.linalg.norm(q[3:7])
        else:
            ppos, prot = (np.zeros(3), np.array([1.0, 0, 0, 0])) if l["parent"] < 0 else world[l["parent"]]
            trans, angle = np.zeros(3), 0.0
            if l["type"] == TYPE_PLANAR:
                trans, angle = np.array([q[0], 0.0, q[1]]), q[2]
            elif l["type"] == TYPE_HINGE:
                angle = q[0]
            elif l["type"] in (TYPE_SLIDE, TYPE_SLIDE2):
                trans = np.asarray(l["axis"], float) * q[0]
            if l["type"] in (TYPE_HINGE2, TYPE_HINGE3):  # stacked hinges: successive rotations about the given axes
                jrot = np.array([1.0, 0, 0, 0])
                for k, a in enumerate(l["axis"]):
                    jrot = quat_mul(jrot, quat_axis_angle(a, q[k]))
            else:
                jrot = quat_axis_angle(l["axis"], angle) if l["axis"] is not None else np.array([1.0, 0, 0, 0])
            lrot = quat_mul(l["quat"], jrot)
            pos = ppos + quat_to_mat(prot) @ (l["pos"] + quat_to_mat(l["quat"]) @ trans)
            rot = quat_mul(prot, lrot)
        world.append((pos, rot))
    return [float((world[li][0] + quat_to_mat(world[li][1]) @ p)[2] - r) for (li, p, r, _fr) in pts_all]


def build_system(model: dict, tunables: dict | None = None) -> dict:
    """Derive masses / inertias (MuJoCo inertiafromgeom) and pack the float32 system table."""
    links = model["links"]
    n = len(links)
    assert n <= MAX_LINKS
    density = model["density"]
    props = [body_inertia(l["geoms"], density) for l in links]
    if model["total_mass"]:  # <compiler settotalmass=...>: uniform rescale of masses and inertias
        scale = model["total_mass"] / sum(p[0] for p in props)
        props = [(m * scale, c, q, w * scale) for m, c, q, w in props]
    tun = dict(model["tunables"])
    if tunables:
        unknown = set(tunables) - set(tun)
        if unknown:
            raise ValueError(f"unknown Brax tunables {sorted(unknown)}; known: {TUNABLE_NAMES}")
        tun.update(tunables)
    t = np.zeros(TABLE_FLOATS, dtype=np.float64)
    names = [l["name"] for l in links]
    # q / qd indexing in link order
    q_idx, qd_idx, qi, qdi = [], [], 0, 0
    for l in links:
        q_idx.append(qi)
        qd_idx.append(qdi)
        nq, nqd = TYPE_DOFS[l["type"]]
        qi += nq
        qdi += nqd
    act_of = {name: i for i, name in enumerate(model["actuator_links"])}
    pts_all = []
    pair_rows = {}
    max_pts = 0
    for i, (l, (m, com, irot, idiag)) in enumerate(zip(links, props)):
        o = OFF_LINKS + LINK_STRIDE * i
        t[o + L_PARENT] = l["parent"]
        t[o + L_TYPE] = l["type"]
        t[o + L_QIDX] = q_idx[i]
        t[o + L_QDIDX] = qd_idx[i]
        t[o + L_TPOS:o + L_TPOS + 3] = l["pos"]
        t[o + L_TROT:o + L_TROT + 4] = l["quat"]
        t[o + L_JPOS:o + L_JPOS + 3] = l["joint_pos"]
        d = OFF_DOF + DOF_STRIDE * i
        t[d + D_ACT1], t[d + D_ACT2] = -1, -1
        t[d + D_SIGN0:d + D_SIGN0 + 3] = 1.0
        if l["type"] in (TYPE_HINGE2, TYPE_HINGE3):
            axes, nd = l["axis"], TYPE_DOFS[l["type"]][0]
            assert len(axes) == nd and len(l["limit"]) == nd and len(l["gear"]) == nd
            jrot = frame_with_axes(axes[0], axes[1])
            t[o + L_JROT:o + L_JROT + 4] = jrot
            t[o + L_LIM_LO], t[o + L_LIM_HI] = l["limit"][0]
            t[o + L_GEAR] = l["gear"][0]
            t[o + L_ACT] = act_of.get(f"{l['name']}:0", -1)
            t[d + D_ACT1], t[d + D_GEAR1] = act_of.get(f"{l['name']}:1", -1), l["gear"][1]
            t[d + D_LO1], t[d + D_HI1] = l["limit"][1]
            if nd == 3:
                zf = quat_to_mat(jrot)[:, 2]
                a2 = np.asarray(axes[2], float) / np.linalg.norm(axes[2])
                assert abs(abs(np.dot(zf, a2)) - 1.0) < 1e-12, "third hinge axis must be +-(first x second)"
                t[d + D_SIGN2] = np.sign(np.dot(zf, a2))
                t[d + D_ACT2], t[d + D_GEAR2] = act_of.get(f"{l['name']}:2", -1), l["gear"][2]
                t[d + D_LO2], t[d + D_HI2] = l["limit"][2]
        else:
            t[o + L_JROT:o + L_JROT + 4] = frame_with_x(l["axis"]) if l["axis"] is not None else (1, 0, 0, 0)
            if l["type"] == TYPE_SLIDE2 and l.get("axis2") is not None:
                t[o + L_JROT:o + L_JROT + 4] = frame_with_axes(l["axis"], l["axis2"])
            t[o + L_LIM_LO], t[o + L_LIM_HI] = l["limit"]
            t[o + L_GEAR] = l["gear"]
            t[o + L_ACT] = act_of.get(l["name"], -1)
        t[o + L_COM:o + L_COM + 3] = com
        t[o + L_IROT:o + L_IROT + 4] = irot
        t[o + L_IDIAG:o + L_IDIAG + 3] = idiag
        t[o + L_MASS] = m
        t[o + L_CTRL_LO], t[o + L_CTRL_HI] = l["ctrl_range"]
        pts = contact_points(l["geoms"]) if model.get("contacts", True) else []
        # body-vs-body pairs: one extra candidate row per pair on either link receives that pair's impulse (radius -1:
        # the ground pass leaves such a row inactive)
        for k, (la, _ga, lb, _gb) in enumerate(model.get("pairs", [])):
            for side, name in ((0, la), (1, lb)):
                if name == l["name"]:
                    pair_rows[(k, side)] = len(pts_all) + len(pts)
                    pts.append((np.zeros(3), -1.0, None))
        t[o + L_FIRST_PT] = len(pts_all)
        t[o + L_N_PT] = len(pts)
        max_pts = max(max_pts, len(pts))
        for (p, r, fr) in pts:
            pts_all.append((i, p, r, model["friction"] if fr is None else fr))
    assert len(pts_all) <= MAX_POINTS
    for k, (li, p, r, fr) in enumerate(pts_all):
        o = OFF_POINTS + POINT_STRIDE * k
        t[o:o + 7] = [li, p[0], p[1], p[2], r, fr, model["stock_elasticity"]]
    # Contact schedule (slot 7 of the point rows): the kernel's contact pass k handles the candidates
    # schedule[k * lanes ...]; candidates closest to the ground in the initial pose (feet) come first, so that the
    # later passes hold the ones that rarely touch and a warp skips their impulse arithmetic altogether. Pure
    # scheduling: impulses are still stored and summed per link in candidate order.
    if pts_all:
        clearance = _initial_point_clearance(links, np.asarray(model["init_q"], float), pts_all, q_idx)
        order = sorted(range(len(pts_all)), key=lambda k: (round(clearance[k], 6), k))
        for slot, k in enumerate(order):
            t[OFF_POINTS + POINT_STRIDE * slot + P_SCHED] = k
    ep = model["env_params"]
    t[H_N_LINKS], t[H_N_Q], t[H_N_QD], t[H_N_POINTS] = n, qi, qdi, len(pts_all)
    t[H_N_FRAMES], t[H_DT], t[H_ENV], t[H_N_ACT] = model["n_frames"], model["dt"], model["env"], len(model["actuator_links"])
    for k, name in enumerate(TUNABLE_NAMES):
        t[H_STIFFNESS + k] = tun[name]
    t[H_RESET_NOISE], t[H_CTRL_COST], t[H_HEALTHY_REWARD] = ep["reset_noise"], ep["ctrl_cost"], ep["healthy_reward"]
    t[H_HEALTHY_Z_MIN], t[H_HEALTHY_Z_MAX], t[H_FORWARD_WEIGHT] = ep["z_min"], ep["z_max"], ep["forward_weight"]
    t[H_ANGLE_MIN], t[H_ANGLE_MAX], t[H_EXCLUDE_POS], t[H_QD_CLIP] = ep["angle_min"], ep["angle_max"], ep["exclude_pos"], ep["qd_clip"]
    t[H_TERMINATE], t[H_MAX_CHILD_POINTS] = ep["terminate"], max_pts
    t[H_QD_UNIFORM] = ep.get("qd_uniform", 0.0)  # Hopper / Walker2d draw qd uniformly, the others N(0,1)
    t[H_QD_NOISE] = ep.get("qd_noise", ep["reset_noise"])
    actuated = {k.split(":")[0] for k in act_of}
    ranges = {tuple(l["ctrl_range"]) for l in links if l["name"] in actuated}
    if not ranges:  # a body without actuators (test bodies)
        ranges = {(-1.0, 1.0)}
    assert len(ranges) == 1 and next(iter(ranges))[0] == -next(iter(ranges))[1], "one symmetric ctrl_range per body"
    t[H_ACT_SCALE] = next(iter(ranges))[1]
    if model.get("site"):
        site_link, site_pos = model["site"]
        t[H_SITE_LINK] = names.index(site_link)
        o = OFF_LINKS + LINK_STRIDE * names.index(site_link)
        t[o + L_SITE:o + L_SITE + 3] = site_pos
    t[OFF_INIT_Q:OFF_INIT_Q + qi] = model["init_q"]
    pairs = model.get("pairs", [])
    assert len(pairs) <= MAX_PAIRS
    t[OFF_PAIR + X_N_PAIRS], t[OFF_PAIR + X_PLANE_Z] = len(pairs), model.get("plane_z", 0.0)
    for k, name in enumerate(model.get("obs_links", ())):
        t[OFF_PAIR + X_OBS_LINK0 + k] = names.index(name)
    for k, (la, ga, lb, gb) in enumerate(pairs):
        o = OFF_PAIR + PAIR_HEADER + PAIR_STRIDE * k
        cap, ball = links[names.index(la)]["geoms"][ga], links[names.index(lb)]["geoms"][gb]
        assert cap["type"] == "capsule" and ball["type"] == "sphere"
        t[o + R_LINK_A], t[o + R_ROW_A] = names.index(la), pair_rows[(k, 0)]
        t[o + R_A0:o + R_A0 + 3], t[o + R_A1:o + R_A1 + 3], t[o + R_RADIUS_A] = cap["p0"], cap["p1"], cap["r"]
        t[o + R_LINK_B], t[o + R_ROW_B] = names.index(lb), pair_rows[(k, 1)]
        t[o + R_B0:o + R_B0 + 3], t[o + R_RADIUS_B] = ball["p0"], ball["r"]
    stock_friction = float(np.max([p[3] for p in pts_all])) if pts_all else float(model["friction"])
    return dict(
        name=model["name"], table=t.astype(np.float32), link_names=names, n_links=n, n_q=qi, n_qd=qdi,
        n_points=len(pts_all), n_act=len(model["actuator_links"]), stock_masses=[float(p[0]) for p in props],
        stock_gravity=model["stock_gravity"], stock_friction=stock_friction,
        stock_elasticity=model["stock_elasticity"], stock_ang_damping=model["stock_ang_damping"],
        tunables=tun, obs_dim=model.get("obs_dim", (qi - int(ep["exclude_pos"])) + qdi), state_words=((13 * n + 3) // 4) * 4,
        act_scale=float(np.float32(t[H_ACT_SCALE])),
        dt=model["dt"] * model["n_frames"],
    )


MODELS = {"ant": ant_model, "halfcheetah": halfcheetah_model, "hopper": hopper_model, "walker2d": walker2d_model,
          "inverted_pendulum": inverted_pendulum_model, "inverted_double_pendulum": inverted_double_pendulum_model,
          "reacher": reacher_model, "humanoid": humanoid_model, "humanoidstandup": humanoidstandup_model,
          "pusher": pusher_model}
SYSTEMS = {k: build_system(f()) for k, f in MODELS.items()}
