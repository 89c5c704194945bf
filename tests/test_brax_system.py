"""Brax body models: the geometry restated from the standard MJCF models is pinned by the
reference's own per-link mass defaults (carl_halfcheetah.py:40-57, carl_hopper.py:40-48), and the
packed-table layout is identical in Python, the CUDA header and the oracle."""
import os
import re

import numpy as np
import pytest

from carl_b200.envs import brax_system as bs
from carl_b200.envs.brax import (CARLBraxAnt, CARLBraxHalfcheetah, CARLBraxHopper, CARLBraxHumanoid,
                                 CARLBraxHumanoidStandup, CARLBraxInvertedDoublePendulum, CARLBraxInvertedPendulum,
                                 CARLBraxPusher, CARLBraxReacher, CARLBraxWalker2d)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_halfcheetah_masses_match_carl_defaults():
    s = bs.SYSTEMS["halfcheetah"]
    d = CARLBraxHalfcheetah.get_context_space().get_default_context()
    for name, m in zip(s["link_names"][1:], s["stock_masses"][1:]):
        assert m == pytest.approx(d[f"mass_{name}"], rel=2e-7), name
    assert sum(s["stock_masses"]) == pytest.approx(14.0, rel=1e-12)  # settotalmass
    assert s["stock_masses"][0] == pytest.approx(6.2502092, rel=2e-7)


def test_hopper_masses_match_carl_defaults():
    s = bs.SYSTEMS["hopper"]
    d = CARLBraxHopper.get_context_space().get_default_context()
    for name, m in zip(s["link_names"][1:], s["stock_masses"][1:]):
        assert m == pytest.approx(d[f"mass_{name}"], rel=2e-7), name
    assert s["stock_masses"][0] == pytest.approx(3.6651914, rel=2e-7)  # CARL overrides it with 10


def test_walker2d_masses_match_carl_defaults():
    """carl/envs/brax/carl_walker2d.py:37-57: six leg-link masses pin the capsule geometry."""
    s = bs.SYSTEMS["walker2d"]
    d = CARLBraxWalker2d.get_context_space().get_default_context()
    for name, m in zip(s["link_names"][1:], s["stock_masses"][1:]):
        assert m == pytest.approx(d[f"mass_{name}"], rel=2e-7), name
    assert s["stock_masses"][0] == pytest.approx(3.6651914, rel=2e-7)  # CARL overrides it with 10
    assert (s["n_links"], s["n_q"], s["n_qd"], s["obs_dim"], s["n_act"]) == (7, 9, 9, 17, 6)
    assert s["dt"] == pytest.approx(0.008)


def test_shapes_match_reference_observation_sizes():
    a, h, p = bs.SYSTEMS["ant"], bs.SYSTEMS["halfcheetah"], bs.SYSTEMS["hopper"]
    assert (a["n_links"], a["n_q"], a["n_qd"], a["obs_dim"], a["n_act"]) == (9, 15, 14, 27, 8)  # obs f32[27]: notebook
    assert (h["n_links"], h["n_q"], h["n_qd"], h["obs_dim"], h["n_act"]) == (7, 9, 9, 17, 6)
    assert (p["n_links"], p["n_q"], p["n_qd"], p["obs_dim"], p["n_act"]) == (4, 6, 6, 11, 3)
    assert a["dt"] == pytest.approx(0.05) and h["dt"] == pytest.approx(0.05) and p["dt"] == pytest.approx(0.008)


def test_every_mass_feature_names_a_link():
    for cls, key in ((CARLBraxAnt, "ant"), (CARLBraxHalfcheetah, "halfcheetah"), (CARLBraxHopper, "hopper"),
                     (CARLBraxWalker2d, "walker2d"), (CARLBraxInvertedPendulum, "inverted_pendulum"),
                     (CARLBraxInvertedDoublePendulum, "inverted_double_pendulum"), (CARLBraxReacher, "reacher"),
                     (CARLBraxHumanoid, "humanoid"), (CARLBraxHumanoidStandup, "humanoidstandup"), (CARLBraxPusher, "pusher")):
        links = bs.SYSTEMS[key]["link_names"]
        for f in cls.get_context_features():
            if f.startswith("mass_"):
                assert f[len("mass_"):] in links


def test_reacher_masses_match_carl_defaults():
    """carl/envs/brax/carl_reacher.py:36-41: mass_body0 / mass_body1 pin the capsule + fingertip geometry."""
    s = bs.SYSTEMS["reacher"]
    d = CARLBraxReacher.get_context_space().get_default_context()
    for name in ("body0", "body1"):
        assert s["stock_masses"][s["link_names"].index(name)] == pytest.approx(d[f"mass_{name}"], rel=2e-7)


def test_pendulum_and_reacher_shapes():
    """brax.envs.inverted_pendulum (obs q ++ qd = 4), inverted_double_pendulum (8), reacher (11); the inverted
    pendulum's action space is its ctrl_range +-3 (BraxGymWrapper, carl/envs/brax/wrappers.py:48-50)."""
    ip, idp, r = (bs.SYSTEMS[k] for k in ("inverted_pendulum", "inverted_double_pendulum", "reacher"))
    assert (ip["n_links"], ip["n_q"], ip["n_qd"], ip["obs_dim"], ip["n_act"], ip["n_points"]) == (2, 2, 2, 4, 1, 0)
    assert (idp["n_links"], idp["n_q"], idp["n_qd"], idp["obs_dim"], idp["n_act"], idp["n_points"]) == (3, 3, 3, 8, 1, 0)
    assert (r["n_links"], r["n_q"], r["n_qd"], r["obs_dim"], r["n_act"], r["n_points"]) == (3, 4, 4, 11, 2, 0)
    assert ip["act_scale"] == 3.0 and idp["act_scale"] == 1.0 and r["act_scale"] == 1.0
    # bug-compatible default context of the double pendulum: no `mass_pole2` key (feature NAME is `mass_pole`)
    d = CARLBraxInvertedDoublePendulum.get_default_context()
    assert "mass_pole2" not in d and "mass_pole" in d
    assert list(CARLBraxInvertedPendulum.get_context_features()) == [
        "gravity", "friction", "elasticity", "mass_cart", "mass_pole", "ang_damping", "viscosity"]


def test_inertia_from_geom_known_values():
    m, com, rot, diag = bs.body_inertia([bs.sphere((0, 0, 0), 0.25)], 5.0)
    assert m == pytest.approx(5 * 4 / 3 * np.pi * 0.25**3) and np.allclose(diag, 0.4 * m * 0.25**2)
    m, com, rot, diag = bs.body_inertia([bs.capsule((0, 0, 0), (0, 0, -0.45), 0.05)], 1000.0)
    assert m == pytest.approx(4.0578905, rel=2e-7) and np.allclose(com, (0, 0, -0.225))
    assert diag[2] < diag[0] and diag[0] == pytest.approx(diag[1])


def _enum_values(text, enum_name):
    body = re.search(r"enum\s+" + enum_name + r"\s*\{(.*?)\};", text, flags=re.S).group(1)
    vals, cur = {}, -1
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = [x.strip() for x in item.split("=")]
            cur = int(v)
        else:
            k, cur = item, cur + 1
        vals[k] = cur
    return vals


def test_table_layout_identical_in_python_and_cuda_header():
    text = open(os.path.join(ROOT, "carl_b200", "csrc", "physics_brax.h")).read()
    for name in ("MAX_LINKS", "MAX_POINTS", "MAX_Q", "HEADER", "LINK_STRIDE", "POINT_STRIDE", "DOF_STRIDE", "MAX_PAIRS",
                 "PAIR_STRIDE", "PAIR_HEADER"):
        assert int(re.search(rf"constexpr int {name} = (\d+);", text).group(1)) == getattr(bs, name)
    hdr = _enum_values(text, "Hdr")
    for k, v in hdr.items():
        assert getattr(bs, k) == v, k
    ls = _enum_values(text, "LinkSlot")
    for k, v in ls.items():
        assert getattr(bs, k) == v, k
    for enum in ("DofSlot", "LinkType", "EnvId", "PairHdr", "PairSlot"):
        for k, v in _enum_values(text, enum).items():
            assert getattr(bs, k) == v, k
    assert bs.OFF_DOF == bs.OFF_INIT_Q + bs.MAX_Q and bs.OFF_PAIR == bs.OFF_DOF + bs.DOF_STRIDE * bs.MAX_LINKS
    assert bs.TABLE_FLOATS == bs.OFF_PAIR + bs.PAIR_HEADER + bs.PAIR_STRIDE * bs.MAX_PAIRS
    assert bs.TABLE_FLOATS * 4 % 16 == 0  # TMA bulk copies move multiples of 16 bytes


def test_tables_are_trees_with_parents_first():
    for s in bs.SYSTEMS.values():
        t = s["table"]
        for l in range(s["n_links"]):
            parent = int(t[bs.OFF_LINKS + bs.LINK_STRIDE * l + bs.L_PARENT])
            assert parent < l
        pts = int(t[bs.H_N_POINTS])
        assert pts <= bs.MAX_POINTS
        firsts = [int(t[bs.OFF_LINKS + bs.LINK_STRIDE * l + bs.L_FIRST_PT]) for l in range(s["n_links"])]
        counts = [int(t[bs.OFF_LINKS + bs.LINK_STRIDE * l + bs.L_N_PT]) for l in range(s["n_links"])]
        assert firsts == list(np.cumsum([0] + counts[:-1])) and sum(counts) == pts


def test_tunable_override_and_validation():
    s = bs.build_system(bs.ant_model(), {"constraint_stiffness": 1234.0})
    assert s["table"][bs.H_STIFFNESS] == 1234.0
    with pytest.raises(ValueError):
        bs.build_system(bs.ant_model(), {"nope": 1.0})


def test_contact_schedule_is_a_permutation_with_the_feet_first():
    """Slot 7 of the point rows: the order in which the kernel's contact passes visit the candidates (pure
    scheduling -- impulses are stored and summed per link in candidate order)."""
    for name, s in bs.SYSTEMS.items():
        t, P = s["table"], s["n_points"]
        sched = [int(t[bs.OFF_POINTS + bs.POINT_STRIDE * k + bs.P_SCHED]) for k in range(P)]
        assert sorted(sched) == list(range(P)), name
    a = bs.SYSTEMS["ant"]
    t = a["table"]
    first = [a["link_names"][int(t[bs.OFF_POINTS + bs.POINT_STRIDE * int(t[bs.OFF_POINTS + bs.POINT_STRIDE * k + bs.P_SCHED])])]
             for k in range(4)]
    assert sorted(first) == ["ankle_1", "ankle_2", "ankle_3", "ankle_4"]  # the four feet lead pass 0
    h = bs.SYSTEMS["hopper"]
    t = h["table"]
    assert h["link_names"][int(t[bs.OFF_POINTS + bs.POINT_STRIDE * int(t[bs.OFF_POINTS + bs.P_SCHED])])] == "foot"


def test_humanoid_geometry_matches_carl_masses():
    """The humanoid's per-link geoms reproduce every MJCF-derived mass default of
    carl/envs/brax/carl_humanoid.py:41-75 to 7 digits (``mass_torso = 10`` there is a placeholder, like the Ant's)."""
    want = {"lwaist": 2.2619467, "pelvis": 6.6161942, "right_thigh": 4.751751, "right_shin": 4.522842,
            "left_thigh": 4.751751, "left_shin": 4.522842, "right_upper_arm": 1.6610805, "right_lower_arm": 1.2295402,
            "left_upper_arm": 1.6610805, "left_lower_arm": 1.2295402}
    geo = bs.humanoid_geometry()
    assert list(geo) == ["torso"] + list(want)
    for name, m in want.items():
        assert bs.body_inertia(geo[name], 1000.0)[0] == pytest.approx(m, rel=2e-7), name
    assert bs.body_inertia(geo["torso"], 1000.0)[0] == pytest.approx(8.907463, rel=1e-6)  # the MJCF's own torso mass


def test_humanoid_tables():
    """carl/envs/brax/carl_humanoid.py / carl_humanoidstandup.py: 11 links, q 24 / qd 23, 17 actuators in the MJCF
    order, the 244-entry observation of brax.envs.humanoid; the masses of both systems are CARL's defaults; the
    stacked-hinge rows carry orthogonal right-handed joint frames, and the hips' third coordinate runs against z."""
    for key, cls in (("humanoid", CARLBraxHumanoid), ("humanoidstandup", CARLBraxHumanoidStandup)):
        s = bs.SYSTEMS[key]
        assert (s["n_links"], s["n_q"], s["n_qd"], s["obs_dim"], s["n_act"], s["n_points"]) == (11, 24, 23, 244, 17, 29)
        assert s["act_scale"] == pytest.approx(0.4) and s["dt"] == pytest.approx(0.015)
        d = cls.get_context_space().get_default_context()
        for name, m in zip(s["link_names"][1:], s["stock_masses"][1:]):
            assert m == pytest.approx(d[f"mass_{name}"], rel=2e-7), name
        t = s["table"]
        acts = []
        for l, name in enumerate(s["link_names"]):
            o, dr = bs.OFF_LINKS + bs.LINK_STRIDE * l, bs.OFF_DOF + bs.DOF_STRIDE * l
            typ = int(t[o + bs.L_TYPE])
            nd = 0 if typ == bs.TYPE_FREE else bs.TYPE_DOFS[typ][1]
            acts += [int(v) for v in (t[o + bs.L_ACT], t[dr + bs.D_ACT1], t[dr + bs.D_ACT2])[:nd]]
            q = t[o + bs.L_JROT:o + bs.L_JROT + 4].astype(np.float64)
            assert abs(np.linalg.norm(q) - 1.0) < 1e-6
            assert np.linalg.det(bs.quat_to_mat(q)) == pytest.approx(1.0, abs=1e-6)
            if name.endswith("thigh"):
                assert typ == bs.TYPE_HINGE3 and t[dr + bs.D_SIGN2] == -1.0
        assert sorted(acts) == list(range(17))
    # standing vs lying: the standup system is the same body expressed in frames turned by -90 deg about y
    up, lying = bs.SYSTEMS["humanoid"]["table"], bs.SYSTEMS["humanoidstandup"]["table"]
    head_up = up[bs.OFF_POINTS + bs.POINT_STRIDE * 2 + 1:bs.OFF_POINTS + bs.POINT_STRIDE * 2 + 4]
    head_ly = lying[bs.OFF_POINTS + bs.POINT_STRIDE * 2 + 1:bs.OFF_POINTS + bs.POINT_STRIDE * 2 + 4]
    np.testing.assert_allclose(head_up, (0, 0, 0.19), atol=1e-7)
    np.testing.assert_allclose(head_ly, (-0.19, 0, 0), atol=1e-7)
    assert lying[bs.OFF_INIT_Q + 2] == pytest.approx(0.105)


def test_pusher_table():
    """carl/envs/brax/carl_pusher.py:36-84: all eight mass defaults (seven arm links + the pushed object) pin the
    restated pusher.xml geometry to 7 digits -- the object's 1.8325957e-3 kg is a ball of radius 0.05 at density 3.5;
    shapes of brax.envs.pusher (obs 23 = q[:7], qd[:7], three centres of mass; 7 actuators with ctrl range +-2); the
    gripper's three capsules are paired with the ball, and every pair owns one candidate row on either link."""
    s = bs.SYSTEMS["pusher"]
    d = CARLBraxPusher.get_context_space().get_default_context()
    for name, m in zip(s["link_names"][:8], s["stock_masses"][:8]):
        assert m == pytest.approx(d[f"mass_{name}"], rel=2e-7), name
    assert (s["n_links"], s["n_q"], s["n_qd"], s["obs_dim"], s["n_act"]) == (9, 11, 11, 23, 7)
    assert s["act_scale"] == 2.0 and s["dt"] == pytest.approx(0.05)
    t = s["table"]
    assert int(t[bs.OFF_PAIR + bs.X_N_PAIRS]) == 3 and t[bs.OFF_PAIR + bs.X_PLANE_Z] == pytest.approx(0.325)
    assert [s["link_names"][int(t[bs.OFF_PAIR + bs.X_OBS_LINK0 + k])] for k in range(3)] == ["r_wrist_flex_link", "object", "goal"]
    rows = set()
    for k in range(3):
        o = bs.OFF_PAIR + bs.PAIR_HEADER + bs.PAIR_STRIDE * k
        la, lb = int(t[o + bs.R_LINK_A]), int(t[o + bs.R_LINK_B])
        assert (s["link_names"][la], s["link_names"][lb]) == ("r_wrist_roll_link", "object")
        for link, row in ((la, int(t[o + bs.R_ROW_A])), (lb, int(t[o + bs.R_ROW_B]))):
            lo = int(t[bs.OFF_LINKS + bs.LINK_STRIDE * link + bs.L_FIRST_PT])
            assert lo <= row < lo + int(t[bs.OFF_LINKS + bs.LINK_STRIDE * link + bs.L_N_PT])
            assert t[bs.OFF_POINTS + bs.POINT_STRIDE * row + 4] == -1.0  # an impulse-only row: no ground candidate
            rows.add(row)
        assert t[o + bs.R_RADIUS_A] == pytest.approx(0.02) and t[o + bs.R_RADIUS_B] == pytest.approx(0.05)
    assert len(rows) == 6
    # slides of the object and the goal: first coordinate along y, second along x (MJCF order obj_slidey, obj_slidex)
    for name in ("object", "goal"):
        o = bs.OFF_LINKS + bs.LINK_STRIDE * s["link_names"].index(name)
        m = bs.quat_to_mat(t[o + bs.L_JROT:o + bs.L_JROT + 4].astype(np.float64))
        np.testing.assert_allclose(m[:, 0], (0, 1, 0), atol=1e-6)
        np.testing.assert_allclose(m[:, 1], (1, 0, 0), atol=1e-6)


def _axis_rot(axis, angle):
    """Rodrigues rotation matrix about a (not necessarily unit) axis."""
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * k + (1 - np.cos(angle)) * (k @ k)


def _model_forward_kinematics(model, q):
    """Link frames (origin, rotation matrix) straight from the MODEL description -- MJCF body positions / quats, joint
    positions and the RAW joint axes, rotations composed in the order the MJCF stacks them -- in float64 numpy. It shares
    nothing with the packed table (joint-frame quaternions, per-dof signs, Euler decomposition) nor with the C / CUDA
    kinematics that consume it."""
    frames, qi = [], 0
    for l in model["links"]:
        typ = l["type"]
        if typ == bs.TYPE_FREE:
            quat = q[qi + 3:qi + 7] / np.linalg.norm(q[qi + 3:qi + 7])
            frames.append((q[qi:qi + 3].astype(np.float64), bs.quat_to_mat(quat)))
            qi += 7
            continue
        ppos, prot = (np.zeros(3), np.eye(3)) if l["parent"] < 0 else frames[l["parent"]]
        body_rot = bs.quat_to_mat(l["quat"])
        trans, rot = np.zeros(3), np.eye(3)
        if typ == bs.TYPE_HINGE:
            rot = _axis_rot(l["axis"], q[qi]); nd = 1
        elif typ == bs.TYPE_PLANAR:    # slide x, slide z, hinge about the given axis
            trans = np.array([q[qi], 0.0, q[qi + 1]]); rot = _axis_rot(l["axis"], q[qi + 2]); nd = 3
        elif typ == bs.TYPE_SLIDE:
            trans = np.asarray(l["axis"], np.float64) * q[qi]; nd = 1
        elif typ == bs.TYPE_SLIDE2:
            second = l.get("axis2")
            if second is None:
                second = bs.quat_to_mat(bs.frame_with_x(l["axis"]))[:, 1]
            trans = np.asarray(l["axis"], np.float64) * q[qi] + np.asarray(second, np.float64) * q[qi + 1]; nd = 2
        else:                          # stacked hinges: successive rotations about the raw MJCF axes
            nd = len(l["axis"])
            for k in range(nd):
                rot = rot @ _axis_rot(l["axis"][k], q[qi + k])
        jp = np.asarray(l["joint_pos"], np.float64)
        local = trans + (jp - rot @ jp)          # the joint position is the pivot of the rotation
        pos = ppos + prot @ (np.asarray(l["pos"], np.float64) + body_rot @ local)
        frames.append((pos, prot @ body_rot @ rot))
        qi += nd
    return frames


@pytest.mark.parametrize("body", list(bs.MODELS))
def test_table_kinematics_match_the_model_description(body):
    """Independent check of everything the packed table encodes about kinematics (link transforms, joint frames built
    from the MJCF axes, the sign of the hips' third coordinate, slide axes, centres of mass): the oracle's
    pipeline_init on the TABLE against a float64 numpy forward kinematics on the MODEL, for random generalized
    coordinates -- link rotation matrices to 2e-6, centre-of-mass positions to 2e-6."""
    from oracle.brax import OracleBraxEnv
    from tests.brax_util import random_q

    model, sysd = bs.MODELS[body](), bs.SYSTEMS[body]
    n, L = 32, sysd["n_links"]
    rng = np.random.default_rng(7)
    q, qd = random_q(sysd, n, rng, scale=4.0)
    ctx = np.zeros((n, 5 + L), np.float32)
    ctx[:, 5:] = np.asarray(sysd["stock_masses"], np.float32)
    ora = OracleBraxEnv(sysd, ctx, f64=True)
    ora.init_from_q(q, 0 * qd)
    rows = ora.state[:, :13 * L].reshape(n, L, 13)
    coms = [bs.body_inertia(l["geoms"], model["density"])[1] for l in model["links"]]
    for e in range(n):
        frames = _model_forward_kinematics(model, q[e].astype(np.float64))
        for l in range(L):
            pos, rot = frames[l]
            np.testing.assert_allclose(bs.quat_to_mat(rows[e, l, 3:7]), rot, atol=2e-6, err_msg=f"{body} link {l} rotation")
            np.testing.assert_allclose(rows[e, l, :3], pos + rot @ coms[l], atol=2e-6, err_msg=f"{body} link {l} COM")


def _geom_quadrature(g, density, h=0.002):
    """Mass, first and second moments of one geom by brute-force voxel integration (float64)."""
    r = g["r"]
    p0 = np.asarray(g["p0"], np.float64)
    p1 = np.asarray(g.get("p1", g["p0"]), np.float64)
    lo, hi = np.minimum(p0, p1) - r, np.maximum(p0, p1) + r
    axes = [np.arange(lo[k] + h / 2, hi[k], h) for k in range(3)]
    x, y, z = np.meshgrid(*axes, indexing="ij")
    pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    ab = p1 - p0
    t = np.zeros(len(pts)) if not ab.any() else np.clip(((pts - p0) @ ab) / (ab @ ab), 0, 1)
    inside = np.linalg.norm(pts - (p0 + t[:, None] * ab), axis=1) <= r
    pts = pts[inside]
    dm = density * h ** 3
    return dm * len(pts), dm * pts.sum(0), dm * (pts[:, :, None] * pts[:, None, :]).sum(0)


@pytest.mark.parametrize("link", ["torso", "right_lower_arm", "left_thigh"])
def test_inertia_from_geom_matches_brute_force_integration(link):
    """MuJoCo's inertiafromgeom as restated in brax_system.body_inertia (capsule = cylinder + two half spheres, geoms
    of a body add up, overlaps counted twice) against voxel integration of the same solids: mass and centre of mass
    (which place every joint anchor and contact candidate relative to the link's COM state) and the principal
    moments -- independent of the closed forms in the code. Composite and oblique geoms of the humanoid."""
    geoms = bs.humanoid_geometry()[link]
    m, com, irot, idiag = bs.body_inertia(geoms, 1000.0)
    mass, first, second = 0.0, np.zeros(3), np.zeros((3, 3))
    for g in geoms:
        a, b, c = _geom_quadrature(g, 1000.0)
        mass, first, second = mass + a, first + b, second + c
    com_q = first / mass
    # (2 mm voxels: the thin arm capsules, r = 31 mm, integrate to ~1 %; a missing hemisphere would be 10 %)
    assert m == pytest.approx(mass, rel=2e-2)
    np.testing.assert_allclose(com, com_q, atol=1e-3)
    inertia_q = (np.trace(second) * np.eye(3) - second) - mass * (com_q @ com_q * np.eye(3) - np.outer(com_q, com_q))
    np.testing.assert_allclose(np.sort(idiag), np.sort(np.linalg.eigvalsh(inertia_q)), rtol=4e-2)
    # and the principal frame really diagonalises it
    rot = bs.quat_to_mat(irot)
    np.testing.assert_allclose(rot.T @ inertia_q @ rot, np.diag(idiag), atol=4e-2 * idiag.max())


def _advance_q(model, q, qd, h):
    """q moved by h along the generalized velocity qd (free roots: position by the linear velocity, orientation by the
    LOCAL angular velocity, as brax's free joint defines its rates)."""
    out, qi, di = q.astype(np.float64).copy(), 0, 0
    for l in model["links"]:
        if l["type"] == bs.TYPE_FREE:
            out[qi:qi + 3] += h * qd[di:di + 3]
            quat = out[qi + 3:qi + 7] / np.linalg.norm(out[qi + 3:qi + 7])
            w = qd[di + 3:di + 6].astype(np.float64)
            ang = np.linalg.norm(w) * h  # signed: h < 0 turns backwards
            dq = np.array([1.0, 0, 0, 0]) if ang == 0 else np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / np.linalg.norm(w)])
            out[qi + 3:qi + 7] = bs.quat_mul(quat, dq)
            qi, di = qi + 7, di + 6
        else:
            nd = bs.TYPE_DOFS[l["type"]][0]
            out[qi:qi + nd] += h * qd[di:di + nd]
            qi, di = qi + nd, di + nd
    return out


@pytest.mark.parametrize("body", list(bs.MODELS))
def test_table_velocity_kinematics_match_finite_differences(body):
    """The velocity half of pipeline_init (link COM velocities and angular velocities from qd: parent chains, joint
    pivots, the stacked hinges' rate composition e_x, Rx e_y, Rx Ry e_z) against central finite differences of the
    model-level forward kinematics above -- independent of the table and of the C / CUDA code."""
    from oracle.brax import OracleBraxEnv
    from tests.brax_util import random_q

    model, sysd = bs.MODELS[body](), bs.SYSTEMS[body]
    n, L = 8, sysd["n_links"]
    rng = np.random.default_rng(8)
    q, qd = random_q(sysd, n, rng, scale=4.0)
    qd = (qd * 5).astype(np.float32)
    ctx = np.zeros((n, 5 + L), np.float32)
    ctx[:, 5:] = np.asarray(sysd["stock_masses"], np.float32)
    ora = OracleBraxEnv(sysd, ctx, f64=True)
    ora.init_from_q(q, qd)
    rows = ora.state[:, :13 * L].reshape(n, L, 13)
    coms = [bs.body_inertia(l["geoms"], model["density"])[1] for l in model["links"]]
    h = 1e-5
    for e in range(n):
        qe = q[e].astype(np.float64)
        if model["links"][0]["type"] == bs.TYPE_FREE:  # pipeline_init normalises the root quaternion
            qe[3:7] /= np.linalg.norm(qe[3:7])
        fp = _model_forward_kinematics(model, _advance_q(model, qe, qd[e], +h))
        fm = _model_forward_kinematics(model, _advance_q(model, qe, qd[e], -h))
        for l in range(L):
            v = ((fp[l][0] + fp[l][1] @ coms[l]) - (fm[l][0] + fm[l][1] @ coms[l])) / (2 * h)
            wx = (fp[l][1] @ fm[l][1].T - fm[l][1] @ fp[l][1].T) / (4 * h)
            w = np.array([wx[2, 1], wx[0, 2], wx[1, 0]])
            np.testing.assert_allclose(rows[e, l, 7:10], v, atol=2e-5, err_msg=f"{body} link {l} COM velocity")
            np.testing.assert_allclose(rows[e, l, 10:13], w, atol=2e-5, err_msg=f"{body} link {l} angular velocity")
