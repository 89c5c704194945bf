// Brax spring-pipeline physics of the batched-step engine: per-link / per-joint / per-contact
// building blocks shared by the warp-per-env CUDA kernels (brax.cu) and the g++-compiled test
// shim (tests/hostcheck).
//
// What this restates: the path `CARLEnv.step` -> `BraxGymWrapper.step` (carl/envs/brax/
// wrappers.py:62-67,74-78) -> brax.envs.{ant,half_cheetah,hopper}.step -> PipelineEnv.pipeline_step
// = n_frames x brax.spring.pipeline.step (maximal coordinates: joint spring/damper forces ->
// semi-implicit velocity update -> ground-contact impulses -> pose integration), plus
// pipeline_init (forward kinematics) and kinematics.inverse (q, qd for the observation).
// brax==0.12.1 (pyproject.toml:62-63) is NOT vendored in the reference tree: the algorithm is
// restated from the published Brax v2 sources (DESIGN.md §brax lists every
// assumption); all system constants come from the table built in carl_b200/envs/brax_system.py.
//
// Conventions: quaternions are (w, x, y, z); per-link state is kept in the centre-of-mass frame
// (pos = world COM, rot = link-frame orientation, vel = COM linear velocity, ang = world angular
// velocity), as Brax integrates `x_i, xd_i`. float32 throughout (the reference's JAX pipeline is
// float32).
#pragma once
#include <math.h>
#include <stdint.h>

#include "rng.h"

namespace carlb {
namespace brax {

// ---- packed system table layout (mirrors carl_b200/envs/brax_system.py) ---------------------
constexpr int MAX_LINKS = 12;
constexpr int MAX_POINTS = 32;
constexpr int MAX_Q = 24;
constexpr int HEADER = 32;
constexpr int LINK_STRIDE = 40;
constexpr int POINT_STRIDE = 8;
constexpr int P_SCHED = 7;  // point row slot 7: contact schedule (candidate handled in this slot of the contact passes)
constexpr int OFF_LINKS = HEADER;
constexpr int OFF_POINTS = OFF_LINKS + LINK_STRIDE * MAX_LINKS;
constexpr int OFF_INIT_Q = OFF_POINTS + POINT_STRIDE * MAX_POINTS;
constexpr int DOF_STRIDE = 16;  // per-link rows of the stacked (2- / 3-dof) revolute joints: what dofs 1 and 2 need
constexpr int OFF_DOF = OFF_INIT_Q + MAX_Q;
// body-vs-body contact pairs (pusher) + what the pusher's env layer reads: header, then MAX_PAIRS rows
constexpr int MAX_PAIRS = 4;
constexpr int PAIR_STRIDE = 16;
constexpr int PAIR_HEADER = 8;
constexpr int OFF_PAIR = OFF_DOF + DOF_STRIDE * MAX_LINKS;
constexpr int TABLE_FLOATS = OFF_PAIR + PAIR_HEADER + PAIR_STRIDE * MAX_PAIRS;
constexpr int MAX_OBS_SMALL = 64;   // observation capacity of the per-env scratch: every body but the humanoids
constexpr int MAX_OBS_LARGE = 244;  // humanoid / humanoidstandup (brax.envs.humanoid._get_obs)
constexpr int MAX_ACT = 20;

// H_SITE_LINK: link carrying the body-fixed point the env layer reads (pendulum tip / reacher fingertip, L_SITE);
// H_QD_NOISE: scale of the reset noise on qd (H_RESET_NOISE is the one on q); H_ACT_SCALE: action-space half width
// (= actuator ctrl_range: 1 except the inverted pendulum's 3)
enum Hdr {
  H_N_LINKS = 0, H_N_Q, H_N_QD, H_N_POINTS, H_N_FRAMES, H_DT, H_ENV, H_N_ACT,
  H_STIFFNESS, H_VEL_DAMPING_C, H_LIMIT_STIFFNESS, H_ANG_DAMPING_C, H_BAUMGARTE, H_VEL_DAMPING, H_MASS_SCALE,
  H_INERTIA_SCALE,
  H_RESET_NOISE, H_CTRL_COST, H_HEALTHY_REWARD, H_HEALTHY_Z_MIN, H_HEALTHY_Z_MAX, H_FORWARD_WEIGHT, H_ANGLE_MIN,
  H_ANGLE_MAX, H_EXCLUDE_POS, H_QD_CLIP, H_TERMINATE, H_MAX_CHILD_POINTS, H_QD_UNIFORM,
  H_SITE_LINK, H_QD_NOISE, H_ACT_SCALE
};
static_assert(H_ACT_SCALE < HEADER, "header slots exhausted");
enum LinkSlot {
  L_PARENT = 0, L_TYPE, L_QIDX, L_QDIDX, L_TPOS = 4, L_TROT = 7, L_JPOS = 11, L_JROT = 14, L_LIM_LO = 18, L_LIM_HI = 19,
  L_COM = 20, L_IROT = 23, L_IDIAG = 27, L_MASS = 30, L_GEAR = 31, L_ACT = 32, L_CTRL_LO = 33, L_CTRL_HI = 34,
  L_FIRST_PT = 35, L_N_PT = 36, L_SITE = 37
};
// TYPE_SLIDE: one prismatic dof along the joint x axis (the cart of the inverted pendulums);
// TYPE_SLIDE2: two prismatic dofs along the joint x and y axes (the reacher's target body)
// TYPE_HINGE2 / TYPE_HINGE3: two / three stacked revolute dofs (the humanoid's abdomen, shoulders / hips): the joint
// rotation is Rx(a0) Ry(a1) Rz(a2) in the joint frame, whose x and y axes are the first two MJCF axes
enum LinkType {
  TYPE_FREE = 0, TYPE_HINGE = 1, TYPE_SLIDE = 2, TYPE_PLANAR = 3, TYPE_SLIDE2 = 4, TYPE_HINGE2 = 5, TYPE_HINGE3 = 6
};
enum EnvId {
  ENV_ANT = 0, ENV_HALFCHEETAH = 1, ENV_HOPPER = 2, ENV_WALKER2D = 3, ENV_INVERTED_PENDULUM = 4,
  ENV_INVERTED_DOUBLE_PENDULUM = 5, ENV_REACHER = 6, ENV_HUMANOID = 7, ENV_HUMANOIDSTANDUP = 8, ENV_PUSHER = 9
};
// pair-region header: pair count, height of the contact plane in the MJCF's world (the engine's plane is z = 0), the
// three links whose centres of mass the pusher observes (wrist-flex link, object, goal)
enum PairHdr { X_N_PAIRS = 0, X_PLANE_Z, X_OBS_LINK0, X_OBS_LINK1, X_OBS_LINK2 };
// pair row: capsule side A (link, candidate row that receives its impulse, segment end points, radius), sphere side B
enum PairSlot { R_LINK_A = 0, R_ROW_A = 1, R_A0 = 2, R_A1 = 5, R_RADIUS_A = 8, R_LINK_B = 9, R_ROW_B = 10, R_B0 = 11, R_RADIUS_B = 14 };
// the bodies of the MODE_SPECIAL kernels (slide joints, their own observation / outcome layers)
CARLB_HD bool is_special_env(int kind) { return (kind >= ENV_INVERTED_PENDULUM && kind <= ENV_REACHER) || kind == ENV_PUSHER; }
// dof rows (OFF_DOF + DOF_STRIDE * link): actuator index / gear / range of dofs 1 and 2 (dof 0 lives in the link row),
// then the sign of each dof's coordinate against the right-handed joint frame (-1 where the MJCF axis is -z)
enum DofSlot { D_ACT1 = 0, D_ACT2, D_GEAR1, D_GEAR2, D_LO1, D_HI1, D_LO2, D_HI2, D_SIGN0, D_SIGN1, D_SIGN2 };
// Kernel flavours (template parameter of the step / reset kernels): which joint types and env layers are compiled in
enum BodyMode { MODE_LOCO = 0, MODE_SPECIAL = 1, MODE_HUMANOID = 2, MODE_PUSHER = 3 };  // PUSHER: SPECIAL + contact pairs
// joint coordinates per link type (free roots are handled separately: 7 / 6)
CARLB_HD int type_ndof(int type) {
  return (type == TYPE_PLANAR || type == TYPE_HINGE3) ? 3 : ((type == TYPE_SLIDE2 || type == TYPE_HINGE2) ? 2 : 1);
}
// per-env context rows: gravity, friction, elasticity, ang_damping, joint-stiffness scale (the
// legacy `joint_stiffness` feature of CARL's docs mapped onto the spring constraint stiffness,
// 1 = stock), then one mass per link
enum CtxRow { C_GRAVITY = 0, C_FRICTION, C_ELASTICITY, C_ANG_DAMPING, C_STIFFNESS_SCALE, C_MASS0 };
constexpr int LINK_WORDS = 13;  // pos3 rot4 vel3 ang3

// ---- small vector algebra -------------------------------------------------------------------
struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };
CARLB_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CARLB_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
CARLB_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
CARLB_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
CARLB_HD V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
CARLB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CARLB_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CARLB_HD float norm(V3 a) { return sqrtf(dot(a, a)); }
CARLB_HD Q4 q4(float w, float x, float y, float z) { Q4 r; r.w = w; r.x = x; r.y = y; r.z = z; return r; }
CARLB_HD Q4 qmul(Q4 a, Q4 b) {
  return q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
CARLB_HD Q4 qconj(Q4 a) { return q4(a.w, -a.x, -a.y, -a.z); }
CARLB_HD Q4 qnormalize(Q4 a) {
  const float n = sqrtf(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z);
  const float inv = 1.0f / n;
  return q4(a.w * inv, a.x * inv, a.y * inv, a.z * inv);
}
// brax.math.rotate: r = 2 (u.v) u + (s^2 - u.u) v + 2 s (u x v)
CARLB_HD V3 rotate(V3 v, Q4 q) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = q.w;
  return 2.0f * dot(u, v) * u + (s * s - dot(u, u)) * v + 2.0f * s * cross(u, v);
}
CARLB_HD V3 inv_rotate(V3 v, Q4 q) { return rotate(v, qconj(q)); }
// rotate((1,0,0), q) and rotate((0,1,0), q) with the multiplications by the literal 0 / 1 components carried out
// by hand (IEEE arithmetic cannot fold x * 0 or x + 0 for the compiler): every surviving term is the one the general
// formula produces, so the results are identical up to the sign of an exact zero.
CARLB_HD V3 rotate_ex(Q4 q) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = q.w, c = s * s - dot(u, u), d = 2.0f * q.x, t = 2.0f * s;
  return v3(d * u.x + c, d * u.y + t * u.z, d * u.z + t * (0.0f - u.y));
}
CARLB_HD V3 rotate_ey(Q4 q) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = q.w, c = s * s - dot(u, u), d = 2.0f * q.y, t = 2.0f * s;
  return v3(d * u.x + t * (0.0f - u.z), d * u.y + c, d * u.z + t * u.x);
}
CARLB_HD V3 rotate_ez(Q4 q) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = q.w, c = s * s - dot(u, u), d = 2.0f * q.z, t = 2.0f * s;
  return v3(d * u.x + t * u.y, d * u.y + t * (0.0f - u.x), d * u.z + c);
}
CARLB_HD Q4 quat_axis_angle(V3 axis, float angle) {
  const float h = 0.5f * angle;
  const float s = sinf(h);
  return q4(cosf(h), axis.x * s, axis.y * s, axis.z * s);
}
CARLB_HD V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
CARLB_HD Q4 ld4(const float* p) { return q4(p[0], p[1], p[2], p[3]); }

struct LinkState {
  V3 pos;  // world COM position
  Q4 rot;  // link frame orientation
  V3 vel;  // COM linear velocity (world)
  V3 ang;  // angular velocity (world)
};
struct Wrench {
  V3 f;  // force (world)
  V3 t;  // torque about the link COM (world)
};

CARLB_HD const float* link_tab(const float* sys, int l) { return sys + OFF_LINKS + LINK_STRIDE * l; }
CARLB_HD const float* point_tab(const float* sys, int p) { return sys + OFF_POINTS + POINT_STRIDE * p; }
CARLB_HD const float* dof_tab(const float* sys, int l) { return sys + OFF_DOF + DOF_STRIDE * l; }
CARLB_HD const float* pair_tab(const float* sys, int k) { return sys + OFF_PAIR + PAIR_HEADER + PAIR_STRIDE * k; }

// link-frame origin in the world (Brax `x.pos`) from the COM state
CARLB_HD V3 link_origin(const LinkState& s, const float* lt) { return s.pos - rotate(ld3(lt + L_COM), s.rot); }
CARLB_HD V3 origin_velocity(const LinkState& s, const float* lt) {
  return s.vel - cross(s.ang, rotate(ld3(lt + L_COM), s.rot));
}

// effective mass / inverse inertia of the spring backend: mass^(1 - spring_mass_scale),
// diag(I)^(1 - spring_inertia_scale) in the principal frame (brax.spring com.inv_inertia)
CARLB_HD float eff_mass(float mass, const float* sys) { return powf(mass, 1.0f - sys[H_MASS_SCALE]); }
CARLB_HD V3 eff_inv_idiag(const float* lt, const float* sys) {
  const float e = 1.0f - sys[H_INERTIA_SCALE];
  return v3(1.0f / powf(lt[L_IDIAG + 0], e), 1.0f / powf(lt[L_IDIAG + 1], e), 1.0f / powf(lt[L_IDIAG + 2], e));
}
// Loop-invariant per-link constants (hoisted out of the substep loop: powf/expf are ~100
// instructions each and the spring step would otherwise evaluate a dozen of them per substep).
struct LinkConst {
  float inv_mass;   // 1 / mass^(1 - spring_mass_scale)
  V3 inv_idiag;     // 1 / diag(I)^(1 - spring_inertia_scale)
  float vel_decay;  // exp(vel_damping * dt)
  float ang_decay;  // exp(ang_damping * dt)
};
CARLB_HD LinkConst make_link_const(const float* sys, const float* lt, float mass, float ang_damping) {
  LinkConst c;
  c.inv_mass = 1.0f / eff_mass(mass, sys);
  c.inv_idiag = eff_inv_idiag(lt, sys);
  c.vel_decay = expf(sys[H_VEL_DAMPING] * sys[H_DT]);
  c.ang_decay = expf(ang_damping * sys[H_DT]);
  return c;
}

CARLB_HD V3 apply_inv_inertia(V3 v, Q4 rot, const float* lt, V3 inv_idiag) {
  const Q4 r = qmul(rot, ld4(lt + L_IROT));
  V3 w = inv_rotate(v, r);
  w = v3(w.x * inv_idiag.x, w.y * inv_idiag.y, w.z * inv_idiag.z);
  return rotate(w, r);
}

// FAST arithmetic (the FMA-contracted build, `arithmetic="fma"`): mathematically identical reformulations that round
// differently from the restated reference order -- with unit effective inertia (spring_inertia_scale = 1, every
// shipped body) I_eff^-1 v is v itself, so the two rotations through the principal frame are skipped.
template <bool FAST>
CARLB_HD V3 apply_inv_inertia_sel(V3 v, Q4 rot, const float* lt, V3 inv_idiag) {
  if (FAST && inv_idiag.x == 1.0f && inv_idiag.y == 1.0f && inv_idiag.z == 1.0f) return v;
  return apply_inv_inertia(v, rot, lt, inv_idiag);
}

// ---- joints (brax.spring.joints.resolve, one joint) ------------------------------------------
// Output of one joint: force/torque in the WORLD frame acting on the child at its anchor `a_c`
// and the opposite reaction on the parent at its anchor `a_p`; plus the joint coordinate (q, qd)
// of kinematics.inverse for the observation.
struct JointOut {
  Wrench child;   // wrench on the child about its COM
  Wrench parent;  // wrench on the parent about its COM (zero for a world parent)
  float q[3];
  float qd[3];
  V3 origin;      // the child's link-frame origin (x.pos), reused by the contact phase of the same substep
};

// child `c` (table row lt), parent state `p` (ignored when world_parent), parent row plt.
// SLIDES = false compiles the prismatic branches out (the locomotion bodies have none: their kernels
// keep the instruction footprint they had before the pendulum / reacher bodies were added).
// anchor of the joint in the parent link frame: link.transform o joint position (table-only, loop invariant)
CARLB_HD V3 parent_anchor(const float* lt) { return ld3(lt + L_TPOS) + rotate(ld3(lt + L_JPOS), ld4(lt + L_TROT)); }

// Table-only facts of a joint that let joint_resolve skip work whose result is known exactly: bit 0 -- the link
// transform has no rotation (q * (1,0,0,0) == q), bit 1 -- the joint sits at the link origin (rotate(0, q) == 0).
// Both hold for every body built so far; the general path stays for tables that differ.
enum JointFlags { JF_TROT_IDENTITY = 1, JF_JPOS_ZERO = 2 };
CARLB_HD int joint_flags(const float* lt) {
  int f = 0;
  if (lt[L_TROT] == 1.0f && lt[L_TROT + 1] == 0.0f && lt[L_TROT + 2] == 0.0f && lt[L_TROT + 3] == 0.0f) f |= JF_TROT_IDENTITY;
  if (lt[L_JPOS] == 0.0f && lt[L_JPOS + 1] == 0.0f && lt[L_JPOS + 2] == 0.0f) f |= JF_JPOS_ZERO;
  return f;
}

// Stacked hinges (kinematics.axis_angle_ang restated as intrinsic x-y'-z'' Euler angles of the joint rotation
// R = Rx(a0) Ry(a1) Rz(a2)): the angles and the axes the limit / actuator torques and the rates refer to, in the
// joint frame: e_x, the line of nodes Rx(a0) e_y, and the child's z axis R e_z.
struct EulerAxes {
  float ang[3];
  V3 a1, a2;
};
CARLB_HD EulerAxes euler_axes(Q4 jrot) {
  EulerAxes e;
  const V3 xc = rotate_ex(jrot), yc = rotate_ey(jrot), zc = rotate_ez(jrot);
  const float c1 = sqrtf(zc.y * zc.y + zc.z * zc.z);
  const float inv = 1.0f / (1e-10f + c1);
  e.ang[0] = atan2f(0.0f - zc.y, zc.z);
  e.ang[1] = atan2f(zc.x, c1);
  e.ang[2] = atan2f(0.0f - yc.x, xc.x);
  e.a1 = v3(0.0f, zc.z * inv, (0.0f - zc.y) * inv);
  e.a2 = zc;
  return e;
}

// STACKED = true compiles the 2- / 3-dof revolute branches in (humanoid kernels only); `dt` is the link's dof row,
// `tau1` / `tau2` the actuator torques of dofs 1 and 2.
// `origins` (optional): the link-frame origins of the child and of the parent, already evaluated by their owners with
// link_origin() on these very states (the kernel shares them through its scratch: one rotation per link and substep
// instead of three); nullptr = evaluate them here. Same function of the same inputs: identical values.
template <bool SLIDES = true, bool STACKED = false>
CARLB_HD JointOut joint_resolve(const float* sys, const float* lt, const LinkState& c, bool world_parent,
                                const float* plt, const LinkState& p, float tau, float stiffness_scale, V3 anchor_p,
                                int flags, const float* dt = nullptr, float tau1 = 0.0f, float tau2 = 0.0f,
                                const V3* origins = nullptr) {
  JointOut o;
  const int type = (int)lt[L_TYPE];
  const Q4 t_rot = ld4(lt + L_TROT), j_rot = ld4(lt + L_JROT);
  const V3 t_pos = ld3(lt + L_TPOS), j_pos = ld3(lt + L_JPOS);
  // anchors (kinematics.world_to_joint): a_c = x_c o joint ; a_p = x_p o link.transform o joint
  const V3 xc_pos = origins != nullptr ? origins[0] : link_origin(c, lt);
  o.origin = xc_pos;
  V3 ac_pos = xc_pos;
  if (!(flags & JF_JPOS_ZERO)) ac_pos = xc_pos + rotate(j_pos, c.rot);
  const Q4 ac_rot = qmul(c.rot, j_rot);
  V3 ap_pos, xp_pos = v3(0, 0, 0), vp = v3(0, 0, 0), wp = v3(0, 0, 0), pcom = v3(0, 0, 0);
  Q4 xp_rot = q4(1, 0, 0, 0);
  if (!world_parent) {
    xp_pos = origins != nullptr ? origins[1] : link_origin(p, plt);
    xp_rot = p.rot;
    wp = p.ang;
    pcom = p.pos;
  }
  ap_pos = xp_pos + rotate(anchor_p, xp_rot);
  Q4 xpt_rot = xp_rot;
  if (!(flags & JF_TROT_IDENTITY)) xpt_rot = qmul(xp_rot, t_rot);
  const Q4 ap_rot = qmul(xpt_rot, j_rot);
  if (!world_parent) vp = p.vel + cross(p.ang, ap_pos - p.pos);
  const V3 vc = c.vel + cross(c.ang, ac_pos - c.pos);
  // joint-frame offsets and rates
  const V3 jpos = inv_rotate(ac_pos - ap_pos, ap_rot);
  const Q4 jrot = qmul(qconj(ap_rot), ac_rot);
  const V3 jvel = inv_rotate(vc - vp, ap_rot);
  const V3 jang = inv_rotate(c.ang - wp, ap_rot);
  const float k = sys[H_STIFFNESS] * stiffness_scale, cv = sys[H_VEL_DAMPING_C], kl = sys[H_LIMIT_STIFFNESS],
              ca = sys[H_ANG_DAMPING_C];
  const V3 ex = v3(1, 0, 0);
  // hinge angle about the joint x axis
  const V3 yc = rotate_ey(jrot);
  const float psi = atan2f(yc.z, yc.y);
  V3 fv, fa;
  // torque aligning the child's joint axis with the parent's
  const V3 axis_c_x = rotate_ex(jrot);
  fa = k * cross(axis_c_x, ex);
  EulerAxes ea;
  const bool stacked = STACKED && (type == TYPE_HINGE2 || type == TYPE_HINGE3);
  if (stacked) {
    // universal / spherical joint: position spring on the anchor; no axis-alignment torque -- the universal joint
    // instead keeps the child's y axis perpendicular to the parent's x axis (third Euler angle = 0); range limit and
    // actuator torque per dof about its Euler axis; damping on the whole relative rate
    ea = euler_axes(jrot);
    fv = (-k) * jpos - cv * jvel;
    fa = v3(0, 0, 0);
    if (type == TYPE_HINGE2) {
      const float inv = 1.0f / (1e-10f + sqrtf(yc.y * yc.y + yc.z * yc.z));
      fa = fa + k * cross(yc, v3(0.0f, yc.y * inv, yc.z * inv));
    }
    const int nd = type == TYPE_HINGE3 ? 3 : 2;
    for (int d = 0; d < nd; ++d) {
      const float lo = d == 0 ? lt[L_LIM_LO] : (d == 1 ? dt[D_LO1] : dt[D_LO2]);
      const float hi = d == 0 ? lt[L_LIM_HI] : (d == 1 ? dt[D_HI1] : dt[D_HI2]);
      const float sg = dt[D_SIGN0 + d];
      const float coord = sg * ea.ang[d];
      float dang = 0.0f;
      if (coord < lo) dang = lo - coord;
      if (coord > hi) dang = hi - coord;
      const float tq = sg * (kl * dang + (d == 0 ? tau : (d == 1 ? tau1 : tau2)));
      fa = fa + tq * (d == 0 ? ex : (d == 1 ? ea.a1 : ea.a2));
    }
    fa = fa - ca * jang;
  } else if (type == TYPE_PLANAR) {
    // slide-x / slide-z / hinge-y root: only the off-plane offset and off-axis rotation are constrained
    fv = v3(-k * jpos.x - cv * jvel.x, 0.0f, 0.0f);
    fa = fa - ca * v3(0.0f, jang.y, jang.z);
  } else if (SLIDES && (type == TYPE_SLIDE || type == TYPE_SLIDE2)) {
    // prismatic: free along the joint x axis (x and y for SLIDE2), the remaining offsets sprung,
    // every rotation locked (second alignment torque on the y axes); limit + actuator force along x
    const V3 ey = v3(0, 1, 0);
    fa = fa + k * cross(yc, ey);
    fa = fa - ca * jang;
    const float lo = lt[L_LIM_LO], hi = lt[L_LIM_HI];
    float dpos = 0.0f;
    if (jpos.x < lo) dpos = lo - jpos.x;
    if (jpos.x > hi) dpos = hi - jpos.x;
    const float fy = (type == TYPE_SLIDE) ? (-k * jpos.y - cv * jvel.y) : 0.0f;
    fv = v3(kl * dpos + tau, fy, -k * jpos.z - cv * jvel.z);
  } else {
    fv = (-k) * jpos - cv * jvel;
    const float lo = lt[L_LIM_LO], hi = lt[L_LIM_HI];
    float dang = 0.0f;
    if (psi < lo) dang = lo - psi;
    if (psi > hi) dang = hi - psi;
    fa = fa + (kl * dang) * ex;
    fa = fa - ca * jang;
    fa = fa + tau * ex;  // actuator torque about the hinge axis
  }
  const V3 F = rotate(fv, ap_rot), T = rotate(fa, ap_rot);
  o.child.f = F;
  o.child.t = T + cross(ac_pos - c.pos, F);
  o.parent.f = v3(0, 0, 0) - F;
  o.parent.t = (v3(0, 0, 0) - T) - cross(ap_pos - pcom, F);
  // joint coordinates for the observation (kinematics.inverse)
  o.q[0] = psi; o.q[1] = 0; o.q[2] = 0;
  o.qd[0] = jang.x; o.qd[1] = 0; o.qd[2] = 0;
  if (type == TYPE_PLANAR) {
    const V3 vo = origin_velocity(c, lt);
    o.q[0] = xc_pos.x - t_pos.x; o.q[1] = xc_pos.z - t_pos.z; o.q[2] = psi;
    o.qd[0] = vo.x; o.qd[1] = vo.z; o.qd[2] = jang.x;
  } else if (SLIDES && (type == TYPE_SLIDE || type == TYPE_SLIDE2)) {
    o.q[0] = jpos.x; o.q[1] = jpos.y;
    o.qd[0] = jvel.x; o.qd[1] = jvel.y;
  }
  if (stacked) {
    // kinematics.inverse: signed Euler angles; rates = projections of the relative rate on the Euler axes
    o.q[0] = dt[D_SIGN0] * ea.ang[0]; o.q[1] = dt[D_SIGN1] * ea.ang[1]; o.q[2] = dt[D_SIGN2] * ea.ang[2];
    o.qd[0] = dt[D_SIGN0] * jang.x; o.qd[1] = dt[D_SIGN1] * dot(ea.a1, jang); o.qd[2] = dt[D_SIGN2] * dot(ea.a2, jang);
  }
  return o;
}

template <bool SLIDES = true, bool STACKED = false>
CARLB_HD JointOut joint_resolve(const float* sys, const float* lt, const LinkState& c, bool world_parent,
                                const float* plt, const LinkState& p, float tau, float stiffness_scale = 1.0f,
                                const float* dt = nullptr, float tau1 = 0.0f, float tau2 = 0.0f) {
  return joint_resolve<SLIDES, STACKED>(sys, lt, c, world_parent, plt, p, tau, stiffness_scale, parent_anchor(lt),
                                        joint_flags(lt), dt, tau1, tau2);
}

// anchor of the joint in the parent link frame relative to the PARENT'S CENTRE OF MASS (table-only): with it the
// world anchor is p.pos + rotate(., p.rot) -- one rotation instead of link_origin + rotate(anchor)
CARLB_HD V3 parent_anchor_from_com(const float* lt, const float* plt, bool world_parent) {
  const V3 a = parent_anchor(lt);
  return world_parent ? a : a - ld3(plt + L_COM);
}

// FAST arithmetic, revolute joints (hinge, planar root, and -- STACKED -- the 2- / 3-dof hinges): the same spring /
// damper / limit / actuator wrench as joint_resolve, evaluated in the WORLD frame. The positional spring and damper of
// a revolute joint are isotropic, so rotating the anchor offset and rate into the joint frame and the force back out is
// the identity (the planar root keeps only the component along its joint x axis); the axis torques use the world
// images of the two joint x axes. 2 rotations + 3 quaternion products instead of 8 + 3. Mathematically identical,
// rounded differently: not comparable bit for bit with the float32 oracle, held to the float64 yardstick like every
// result of the FMA build (tests/test_brax_parity_gpu.py). Every joint lane of a warp takes this one path.
template <bool STACKED = false>
CARLB_HD JointOut joint_resolve_world(const float* sys, const float* lt, const LinkState& c, bool world_parent,
                                      const LinkState& p, float tau, float stiffness_scale, V3 anchor_pc, int flags,
                                      const float* dt = nullptr, float tau1 = 0.0f, float tau2 = 0.0f) {
  JointOut o;
  const int type = (int)lt[L_TYPE];
  const Q4 j_rot = ld4(lt + L_JROT);
  V3 rc = v3(0, 0, 0) - ld3(lt + L_COM);
  if (!(flags & JF_JPOS_ZERO)) rc = ld3(lt + L_JPOS) + rc;
  const V3 lc = rotate(rc, c.rot);  // child anchor relative to the child's COM, world axes
  const V3 ac_pos = c.pos + lc;
  o.origin = ac_pos;  // (the FAST substeps do not use link origins)
  V3 lp = anchor_pc, wp = v3(0, 0, 0), vp = v3(0, 0, 0), ap_pos = anchor_pc;
  Q4 tj = j_rot;
  if (!(flags & JF_TROT_IDENTITY)) tj = qmul(ld4(lt + L_TROT), j_rot);
  Q4 ap_rot = tj;
  if (!world_parent) {
    lp = rotate(anchor_pc, p.rot);
    ap_pos = p.pos + lp;
    wp = p.ang;
    vp = p.vel + cross(p.ang, lp);
    ap_rot = qmul(p.rot, tj);
  }
  const V3 vc = c.vel + cross(c.ang, lc);
  const Q4 ac_rot = qmul(c.rot, j_rot);
  const Q4 jrot = qmul(qconj(ap_rot), ac_rot);
  const V3 yc = rotate_ey(jrot);
  const float psi = atan2f(yc.z, yc.y);
  const V3 xp_w = rotate_ex(ap_rot), xc_w = rotate_ex(ac_rot);
  const float k = sys[H_STIFFNESS] * stiffness_scale, cv = sys[H_VEL_DAMPING_C], kl = sys[H_LIMIT_STIFFNESS],
              ca = sys[H_ANG_DAMPING_C];
  const V3 d = ac_pos - ap_pos, vrel = vc - vp, wrel = c.ang - wp;
  V3 F = (-k) * d - cv * vrel;
  V3 T;
  if (STACKED && (type == TYPE_HINGE2 || type == TYPE_HINGE3)) {
    // limit / actuator torques about the Euler axes and the universal joint's constraint torque are assembled in the
    // joint frame (as joint_resolve does) and carried out with ONE rotation
    const EulerAxes ea = euler_axes(jrot);
    V3 fa = v3(0, 0, 0);
    if (type == TYPE_HINGE2) {
      const float inv = 1.0f / (1e-10f + sqrtf(yc.y * yc.y + yc.z * yc.z));
      fa = k * cross(yc, v3(0.0f, yc.y * inv, yc.z * inv));
    }
    const int nd = type == TYPE_HINGE3 ? 3 : 2;
    for (int a = 0; a < nd; ++a) {
      const float lo = a == 0 ? lt[L_LIM_LO] : (a == 1 ? dt[D_LO1] : dt[D_LO2]);
      const float hi = a == 0 ? lt[L_LIM_HI] : (a == 1 ? dt[D_HI1] : dt[D_HI2]);
      const float sg = dt[D_SIGN0 + a];
      const float coord = sg * ea.ang[a];
      float dang = 0.0f;
      if (coord < lo) dang = lo - coord;
      if (coord > hi) dang = hi - coord;
      const float tq = sg * (kl * dang + (a == 0 ? tau : (a == 1 ? tau1 : tau2)));
      fa = fa + tq * (a == 0 ? v3(1, 0, 0) : (a == 1 ? ea.a1 : ea.a2));
    }
    T = rotate(fa, ap_rot) - ca * wrel;
  } else if (type == TYPE_PLANAR) {
    // slide-x / slide-z / hinge-y root: only the off-plane offset and the off-axis rotation are constrained
    F = ((-k) * dot(d, xp_w) - cv * dot(vrel, xp_w)) * xp_w;
    T = k * cross(xc_w, xp_w) - ca * (wrel - dot(wrel, xp_w) * xp_w);
  } else {
    const float lo = lt[L_LIM_LO], hi = lt[L_LIM_HI];
    float dang = 0.0f;
    if (psi < lo) dang = lo - psi;
    if (psi > hi) dang = hi - psi;
    T = k * cross(xc_w, xp_w) + (kl * dang + tau) * xp_w - ca * wrel;
  }
  o.child.f = F;
  o.child.t = T + cross(lc, F);
  o.parent.f = v3(0, 0, 0) - F;
  o.parent.t = (v3(0, 0, 0) - T) - cross(lp, F);
  // (the generalized coordinates of the observation always come from joint_resolve)
  o.q[0] = psi; o.q[1] = 0; o.q[2] = 0;
  o.qd[0] = dot(wrel, xp_w); o.qd[1] = 0; o.qd[2] = 0;
  return o;
}

// ---- semi-implicit velocity update (brax.spring.integrator.integrate_xdd) ---------------------
template <bool FAST = false>
CARLB_HD void integrate_xdd(LinkState& s, const Wrench& w, const float* sys, const float* lt, const LinkConst& lc,
                            float gravity) {
  const float dt = sys[H_DT];
  const V3 acc = v3(0, 0, gravity) + w.f * lc.inv_mass;
  const V3 alpha = apply_inv_inertia_sel<FAST>(w.t, s.rot, lt, lc.inv_idiag);
  s.vel = s.vel + acc * dt;
  s.ang = s.ang + alpha * dt;
  s.vel = s.vel * lc.vel_decay;
  s.ang = s.ang * lc.ang_decay;
}

// ---- ground contact of one candidate point (brax.spring.collisions._collide vs the plane z=0) --
struct ContactOut {
  V3 p;        // linear impulse on the link
  V3 t;        // angular impulse about the link COM
  float active;
};

// `origin` is link_origin(s, lt): pose integration is the last phase of a substep, so the value the joint phase
// computed for the link is still exact here and is passed in instead of being recomputed per contact point.
// Height of a link's centre of mass above which contact candidate `pt` cannot touch the ground whatever the link's
// orientation: |candidate - COM| + radius, plus a margin far above float32 rounding. A pure early-out: a candidate
// skipped by this bound would have failed the penetration test.
CARLB_HD float contact_reach(const float* pt, const float* lt) {
  return norm(v3(pt[1], pt[2], pt[3]) - ld3(lt + L_COM)) + pt[4] + 1e-3f;
}

// FAST: the sphere centre is taken from the centre of mass (pos + rotate(candidate - COM)) instead of the link
// origin, which the FAST substeps never evaluate; `origin` is then unused.
template <bool FAST = false>
CARLB_HD ContactOut contact_resolve(const float* sys, const float* pt, const float* lt, const LinkState& s,
                                    const LinkConst& lc, float friction, float elasticity, V3 origin) {
  ContactOut o;
  o.p = v3(0, 0, 0); o.t = v3(0, 0, 0); o.active = 0.0f;
  const float radius = pt[4];
  if (radius < 0.0f) return o;  // a row that only receives the impulse of a body-vs-body pair
  const V3 c = FAST ? s.pos + rotate(v3(pt[1], pt[2], pt[3]) - ld3(lt + L_COM), s.rot)
                    : origin + rotate(v3(pt[1], pt[2], pt[3]), s.rot);  // sphere centre in the world
  const float dist = c.z - radius;                               // signed distance to the plane
  const float penetration = -dist;
  if (!(penetration > 0.0f)) return o;
  // The contact normal of the ground plane is n = (0, 0, 1). Brax's formulas below are written with the
  // products by n's literal 0 / 1 components carried out by hand (the compiler may not fold x * 0 or x + 0 under
  // IEEE rules): dot(n, v) = v.z, cross(r, n) = (r.y, -r.x, 0), s * n = (0, 0, s) -- the same values up to the
  // sign of an exact zero.
  const V3 cpos = v3(c.x, c.y, 0.5f * dist);  // midway between the two surfaces
  const V3 rel_pos = cpos - s.pos;
  const V3 rel_vel = s.vel + cross(s.ang, rel_pos);
  const float normal_vel = rel_vel.z;                                       // dot(n, rel_vel)
  const float inv_m = lc.inv_mass;
  const V3 temp1 = apply_inv_inertia_sel<FAST>(v3(rel_pos.y, 0.0f - rel_pos.x, 0.0f), s.rot, lt, lc.inv_idiag);  // cross(rel_pos, n)
  const float ang = temp1.x * rel_pos.y - temp1.y * rel_pos.x;              // dot(n, cross(temp1, rel_pos))
  const float dt = sys[H_DT];
  const float baumgarte_vel = sys[H_BAUMGARTE] * penetration / dt;
  const float impulse = (-1.0f * (1.0f + elasticity) * normal_vel + baumgarte_vel) / (inv_m + ang);
  const V3 impulse_vec = v3(0.0f, 0.0f, impulse);                           // impulse * n
  // drag from friction, parallel to the surface
  const V3 vel_d = v3(rel_vel.x, rel_vel.y, rel_vel.z - normal_vel);        // rel_vel - normal_vel * n
  const float speed_d = norm(vel_d);
  float impulse_d = speed_d / (inv_m + ang);
  const V3 dir_d = vel_d * (1.0f / (1e-6f + speed_d));
  impulse_d = fminf(impulse_d, friction * impulse);
  const V3 impulse_d_vec = (-impulse_d) * dir_d;
  const bool apply_n = (normal_vel < 0.0f) && (impulse > 0.0f);
  const bool apply_d = apply_n && (speed_d > 0.01f);
  if (!apply_n) return o;
  V3 total = impulse_vec;
  if (apply_d) total = total + impulse_d_vec;
  o.p = total;
  o.t = cross(rel_pos, total);
  o.active = 1.0f;
  return o;
}

// ---- body-vs-body contact of one pair (spring/collisions.py with two dynamic bodies): capsule A vs sphere B --------
// The normal points from B's centre to the closest point of A's segment (it pushes A away from B), the contact point
// lies midway between the two surfaces; the impulse divides by both inverse masses and both angular terms and acts
// with opposite signs on the two links.
struct PairOut {
  V3 p;       // linear impulse on A (B receives -p)
  V3 ta, tb;  // angular impulses about the two centres of mass (tb already carries its sign)
  float active;
};
CARLB_HD PairOut pair_resolve(const float* sys, const float* pr, const float* lta, const float* ltb, const LinkState& sa,
                              const LinkState& sb, const LinkConst& lca, const LinkConst& lcb, float friction, float elasticity) {
  PairOut o;
  o.p = v3(0, 0, 0); o.ta = v3(0, 0, 0); o.tb = v3(0, 0, 0); o.active = 0.0f;
  const V3 oa = link_origin(sa, lta), ob = link_origin(sb, ltb);
  const V3 a0 = oa + rotate(ld3(pr + R_A0), sa.rot), a1 = oa + rotate(ld3(pr + R_A1), sa.rot);
  const V3 cb = ob + rotate(ld3(pr + R_B0), sb.rot);
  const V3 ab = a1 - a0, ac = cb - a0;
  float tt = dot(ac, ab) / dot(ab, ab);
  tt = fminf(fmaxf(tt, 0.0f), 1.0f);
  const V3 cp = a0 + tt * ab;  // closest point of the segment to the sphere's centre
  const V3 dvec = cp - cb;
  const float dist = norm(dvec);
  const float ra_ = pr[R_RADIUS_A], rb_ = pr[R_RADIUS_B];
  const float penetration = ra_ + rb_ - dist;
  if (!(penetration > 0.0f)) return o;
  const V3 n = dvec * (1.0f / (1e-6f + dist));
  const V3 cpos = 0.5f * ((cb + rb_ * n) + (cp - ra_ * n));
  const V3 ra = cpos - sa.pos, rb = cpos - sb.pos;
  const V3 contact_vel = (sa.vel + cross(sa.ang, ra)) - (sb.vel + cross(sb.ang, rb));
  const float normal_vel = dot(n, contact_vel);
  const V3 t1a = apply_inv_inertia(cross(ra, n), sa.rot, lta, lca.inv_idiag);
  const V3 t1b = apply_inv_inertia(cross(rb, n), sb.rot, ltb, lcb.inv_idiag);
  const float ang = dot(n, cross(t1a, ra)) + dot(n, cross(t1b, rb));
  const float denom = lca.inv_mass + lcb.inv_mass + ang;
  const float baumgarte_vel = sys[H_BAUMGARTE] * penetration / sys[H_DT];
  const float impulse = (-1.0f * (1.0f + elasticity) * normal_vel + baumgarte_vel) / denom;
  const V3 vel_d = contact_vel - normal_vel * n;
  const float speed_d = norm(vel_d);
  float impulse_d = speed_d / denom;
  const V3 dir_d = vel_d * (1.0f / (1e-6f + speed_d));
  impulse_d = fminf(impulse_d, friction * impulse);
  const bool apply_n = (normal_vel < 0.0f) && (impulse > 0.0f);
  const bool apply_d = apply_n && (speed_d > 0.01f);
  if (!apply_n) return o;
  V3 total = impulse * n;
  if (apply_d) total = total + (-impulse_d) * dir_d;
  o.p = total;
  o.ta = cross(ra, total);
  o.tb = v3(0, 0, 0) - cross(rb, total);
  o.active = 1.0f;
  return o;
}

// delta-velocity from the link's summed contact impulses, averaged over its active contacts
template <bool FAST = false>
CARLB_HD void integrate_xdv(LinkState& s, V3 p_sum, V3 t_sum, float n_active, const float* lt, const LinkConst& lc) {
  if (!(n_active > 0.0f)) return;
  const float inv_n = 1.0f / n_active;
  s.vel = s.vel + p_sum * (inv_n * lc.inv_mass);
  s.ang = s.ang + apply_inv_inertia_sel<FAST>(t_sum * inv_n, s.rot, lt, lc.inv_idiag);
}

// ---- pose integration (brax.spring.integrator.integrate) --------------------------------------
CARLB_HD void integrate_pose(LinkState& s, float dt) {
  s.pos = s.pos + s.vel * dt;
  // d = qmul((0, w), rot) with the products by the zero scalar part written out (0*x - t == -t, 0*x + t == t)
  const float wx = s.ang.x * 0.5f * dt, wy = s.ang.y * 0.5f * dt, wz = s.ang.z * 0.5f * dt;
  const Q4 r = s.rot;
  const Q4 d = q4((0.0f - wx * r.x) - wy * r.y - wz * r.z, wx * r.w + wy * r.z - wz * r.y, (wy * r.w - wx * r.z) + wz * r.x,
                  wx * r.y - wy * r.x + wz * r.w);
  s.rot = qnormalize(q4(s.rot.w + d.w, s.rot.x + d.x, s.rot.y + d.y, s.rot.z + d.z));
}

// ---- forward kinematics of one link (kinematics.forward + com.from_world), parent first ------
template <bool STACKED = false>
CARLB_HD LinkState forward_link(const float* sys, const float* lt, const float* q, const float* qd, bool world_parent,
                                const float* plt, const LinkState& p, const float* dt = nullptr) {
  const int type = (int)lt[L_TYPE];
  const float* ql = q + (int)lt[L_QIDX];
  const float* qdl = qd + (int)lt[L_QDIDX];
  V3 xpos, xvel, xang;
  Q4 xrot;
  if (type == TYPE_FREE) {
    xpos = v3(ql[0], ql[1], ql[2]);
    xrot = qnormalize(q4(ql[3], ql[4], ql[5], ql[6]));
    xvel = v3(qdl[0], qdl[1], qdl[2]);
    xang = rotate(v3(qdl[3], qdl[4], qdl[5]), xrot);
  } else {
    const Q4 t_rot = ld4(lt + L_TROT), j_rot = ld4(lt + L_JROT);
    const V3 t_pos = ld3(lt + L_TPOS), j_pos = ld3(lt + L_JPOS);
    const V3 axis = rotate(v3(1, 0, 0), j_rot);  // hinge axis in the child link frame
    V3 xp_pos = v3(0, 0, 0), vp = v3(0, 0, 0), wp = v3(0, 0, 0);
    Q4 xp_rot = q4(1, 0, 0, 0);
    if (!world_parent) {
      xp_pos = link_origin(p, plt);
      xp_rot = p.rot;
      vp = origin_velocity(p, plt);
      wp = p.ang;
    }
    float angle, rate;
    V3 trans = v3(0, 0, 0), tvel = v3(0, 0, 0);
    const bool stacked = STACKED && (type == TYPE_HINGE2 || type == TYPE_HINGE3);
    if (stacked) {
      angle = 0.0f;
      rate = 0.0f;
    } else if (type == TYPE_PLANAR) {
      trans = v3(ql[0], 0.0f, ql[1]);
      tvel = v3(qdl[0], 0.0f, qdl[1]);
      angle = ql[2];
      rate = qdl[2];
    } else if (type == TYPE_SLIDE || type == TYPE_SLIDE2) {
      trans = axis * ql[0];
      tvel = axis * qdl[0];
      if (type == TYPE_SLIDE2) {
        const V3 axis_y = rotate(v3(0, 1, 0), j_rot);
        trans = trans + axis_y * ql[1];
        tvel = tvel + axis_y * qdl[1];
      }
      angle = 0.0f;
      rate = 0.0f;
    } else {
      angle = ql[0];
      rate = qdl[0];
    }
    Q4 jrot = quat_axis_angle(axis, angle);
    V3 wj = v3(0, 0, 0);  // stacked hinges: relative angular velocity in the joint frame
    if (stacked) {
      // joint rotation Rx(a0) Ry(a1) Rz(a2) in the joint frame, carried into the link frame by the joint
      // orientation; rates about e_x, Rx(a0) e_y, Rx(a0) Ry(a1) e_z
      Q4 acc = q4(1, 0, 0, 0);
      const int nd = type == TYPE_HINGE3 ? 3 : 2;
      for (int d = 0; d < nd; ++d) {
        const float sg = dt[D_SIGN0 + d];
        const V3 b = v3(d == 0 ? 1.0f : 0.0f, d == 1 ? 1.0f : 0.0f, d == 2 ? 1.0f : 0.0f);
        wj = wj + (sg * qdl[d]) * rotate(b, acc);
        acc = qmul(acc, quat_axis_angle(b, sg * ql[d]));
      }
      jrot = qmul(qmul(j_rot, acc), qconj(j_rot));
    }
    jrot = qnormalize(jrot);
    const V3 jpos = trans + (j_pos - rotate(j_pos, jrot));  // the joint position is the rotation pivot
    const V3 lpos = t_pos + rotate(jpos, t_rot);
    const Q4 lrot = qmul(t_rot, jrot);
    xpos = xp_pos + rotate(lpos, xp_rot);
    xrot = qnormalize(qmul(xp_rot, lrot));
    xvel = vp + cross(wp, xpos - xp_pos) + rotate(tvel, xp_rot);
    xang = wp + rotate(axis * rate, xrot);
    if (stacked) xang = wp + rotate(rotate(wj, j_rot), qmul(xp_rot, t_rot));
    if (j_pos.x != 0.0f || j_pos.y != 0.0f || j_pos.z != 0.0f) {
      // a joint away from the link origin is the PIVOT of the rotation: the origin sits at pivot - R j_pos and is
      // carried around the pivot at -w x (R j_pos), w = the joint's own angular rate in the link-transform frame
      // (bodies whose joints sit at their link origins -- all but the humanoids' waist and knees -- never get here)
      const V3 wl = stacked ? rotate(wj, j_rot) : axis * rate;
      const V3 swing = v3(0, 0, 0) - cross(wl, rotate(j_pos, jrot));
      xvel = xvel + rotate(rotate(swing, t_rot), xp_rot);
    }
  }
  LinkState s;
  const V3 rc = rotate(ld3(lt + L_COM), xrot);
  s.pos = xpos + rc;
  s.rot = xrot;
  s.vel = xvel + cross(xang, rc);
  s.ang = xang;
  return s;
}

// ---- humanoid observation pieces (brax.envs.humanoid._get_obs / _com) -----------------------------------------
// centre of mass of the whole body with the spring backend's effective link masses; `rows` = L link rows of
// LINK_WORDS floats (COM position first), `meff` = effective mass per link. Returns the mass sum.
CARLB_HD float body_com(const float* rows, const float* meff, int L, V3& com) {
  float msum = 0.0f;
  com = v3(0, 0, 0);
  for (int l = 0; l < L; ++l) {
    const float m = meff[l];
    msum += m;
    com = com + m * ld3(rows + l * LINK_WORDS);
  }
  com = v3(com.x / msum, com.y / msum, com.z / msum);
  return msum;
}
// cinert row of one link: its inertia about the body COM in world axes (3x3 row-major) and its mass:
// R diag(I_eff) R^T + m (|p|^2 E - p p^T), R = link rotation o principal frame, p = link COM - body COM
CARLB_HD void link_cinert(const float* sys, const float* lt, const LinkState& s, float m, V3 com, float* out) {
  const V3 p = s.pos - com;
  const Q4 r = qmul(s.rot, ld4(lt + L_IROT));
  const V3 c0 = rotate_ex(r), c1 = rotate_ey(r), c2 = rotate_ez(r);
  const float R[3][3] = {{c0.x, c1.x, c2.x}, {c0.y, c1.y, c2.y}, {c0.z, c1.z, c2.z}};
  const float e = 1.0f - sys[H_INERTIA_SCALE];
  const float ie[3] = {powf(lt[L_IDIAG + 0], e), powf(lt[L_IDIAG + 1], e), powf(lt[L_IDIAG + 2], e)};
  const float pv[3] = {p.x, p.y, p.z};
  const float pp = dot(p, p);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      const float rot_i = (R[a][0] * ie[0] * R[b][0] + R[a][1] * ie[1] * R[b][1]) + R[a][2] * ie[2] * R[b][2];
      const float par = (a == b ? pp : 0.0f) - pv[a] * pv[b];
      out[3 * a + b] = rot_i + m * par;
    }
  out[9] = m;
}
// cvel row of one link: mass-weighted COM velocity and the angular velocity
CARLB_HD void link_cvel(const LinkState& s, float m, float msum, float* out) {
  out[0] = m * s.vel.x / msum; out[1] = m * s.vel.y / msum; out[2] = m * s.vel.z / msum;
  out[3] = s.ang.x; out[4] = s.ang.y; out[5] = s.ang.z;
}
// humanoid observation layout: q[2:] (nq - 2) | qd | cinert 10 L | cvel 6 L | actuator torques (qd layout)
CARLB_HD int humanoid_obs_dim(int nq, int nqd, int L) { return (nq - 2) + nqd + 16 * L + nqd; }

// Observation entry i of the bodies whose obs is not q[exclude:] ++ qd
// (brax.envs.inverted_double_pendulum._get_obs, brax.envs.reacher._get_obs)
CARLB_HD float special_obs_entry(int kind, int i, const float* q, const float* qd, V3 site) {
  if (kind == ENV_INVERTED_DOUBLE_PENDULUM) {  // q[:1], sin(q[1:]), cos(q[1:]), clip(qd, -10, 10)
    if (i == 0) return q[0];
    if (i < 3) return sinf(q[i]);
    if (i < 5) return cosf(q[i - 2]);
    return fminf(fmaxf(qd[i - 5], -10.0f), 10.0f);
  }
  // reacher: cos(theta), sin(theta), q[2:] (target xy), qd[:2], tip - target
  if (i < 2) return cosf(q[i]);
  if (i < 4) return sinf(q[i - 2]);
  if (i < 6) return q[i - 2];
  if (i < 8) return qd[i - 6];
  return i == 8 ? site.x : (i == 9 ? site.y : site.z);
}

// Env layer of the bodies without a locomotion root (brax.envs.inverted_pendulum / inverted_double_pendulum /
// reacher .step); `site` is the pendulum tip in the world, or fingertip - target for the reacher.
CARLB_HD void special_outcome(int kind, float q1, float qd1, float qd2, V3 site, float act_sq_sum, float& reward, bool& done) {
  if (kind == ENV_INVERTED_PENDULUM) {  // reward 1, done = |pole angle| > 0.2
    reward = 1.0f;
    done = fabsf(q1) > 0.2f;
  } else if (kind == ENV_INVERTED_DOUBLE_PENDULUM) {
    const float x = site.x, y = site.z;
    const float dist_penalty = 0.01f * (x * x) + (y - 2.0f) * (y - 2.0f);
    const float vel_penalty = 1e-3f * (qd1 * qd1) + 5e-3f * (qd2 * qd2);
    reward = 10.0f - dist_penalty - vel_penalty;
    done = y <= 1.0f;
  } else if (kind == ENV_REACHER) {  // -|tip - target| - sum(a^2), never done
    reward = (0.0f - norm(site)) + (0.0f - act_sq_sum);
    done = false;
  }
}

// brax.envs.pusher: entry i of the observation (q[:7], qd[:7], centre of mass of the wrist-flex link, the object and
// the goal, heights relative to the MJCF's world) from the env's link rows; and what its reward reads: the distances
// object - fingertip link (x) and object - goal (y) of the state BEFORE the pipeline advances
CARLB_HD float pusher_obs_entry(const float* sys, int i, const float* q, const float* qd, const float* rows) {
  if (i < 7) return q[i];
  if (i < 14) return qd[i - 7];
  const int j = (i - 14) / 3, k = (i - 14) % 3;
  const float v = rows[(int)sys[OFF_PAIR + X_OBS_LINK0 + j] * LINK_WORDS + k];
  return k == 2 ? v - sys[OFF_PAIR + X_PLANE_Z] : v;
}
CARLB_HD V3 pusher_distances(const float* sys, const float* rows) {
  const V3 tip = ld3(rows + (int)sys[OFF_PAIR + X_OBS_LINK0] * LINK_WORDS), obj = ld3(rows + (int)sys[OFF_PAIR + X_OBS_LINK1] * LINK_WORDS),
           goal = ld3(rows + (int)sys[OFF_PAIR + X_OBS_LINK2] * LINK_WORDS);
  return v3(norm(obj - tip), norm(obj - goal), 0.0f);
}
// reward_dist + 0.1 reward_ctrl + 0.5 reward_near, never done (`before` = pusher_distances of the pre-step state)
CARLB_HD void pusher_outcome(const float* sys, V3 before, float act_sq_sum, float& reward, bool& done) {
  reward = ((0.0f - before.y) + sys[H_CTRL_COST] * (0.0f - act_sq_sum)) + 0.5f * (0.0f - before.x);
  done = false;
}

// world position of the env's site: x.take(link).do(Transform(pos=site)); for the reacher minus the target origin
CARLB_HD V3 site_position(const float* sys, const LinkState& site_link, const LinkState& link2) {
  const float* slt = link_tab(sys, (int)sys[H_SITE_LINK]);
  V3 p = link_origin(site_link, slt) + rotate(ld3(slt + L_SITE), site_link.rot);
  if ((int)sys[H_ENV] == ENV_REACHER) p = p - link_origin(link2, link_tab(sys, 2));
  return p;
}

// Reset-noise draws (throughput-mode RNG; the reference's JAX threefry stream is not
// reproducible without JAX): one Philox block per (seed, env, episode, index).
CARLB_HD float reset_uniform(uint64_t seed, uint64_t env, uint32_t episode, uint32_t idx, float lo, float hi) {
  const Philox4 r = philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), episode, 0x52535430u + idx, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
  return lo + (hi - lo) * u32_to_unit_float(r.v[0]);
}
CARLB_HD float reset_normal(uint64_t seed, uint64_t env, uint32_t episode, uint32_t idx) {
  const Philox4 r = philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), episode, 0x4e524d30u + idx, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
  const float u1 = ((float)(r.v[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = u32_to_unit_float(r.v[1]);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

}  // namespace brax
}  // namespace carlb
