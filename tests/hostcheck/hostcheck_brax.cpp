// TEST-ONLY shim: the product's __host__ __device__ Brax physics (carl_b200/csrc/physics_brax.h)
// driven by a serial loop that mirrors the phases of the warp-per-env kernel in brax.cu, compiled
// by g++ so the no-GPU suite can compare it with the oracle. Never loaded by the product.
#include <stdint.h>
#include <string.h>

#include "../../carl_b200/csrc/physics_brax.h"

using namespace carlb;
using namespace carlb::brax;

static LinkState rd(const float* p) {
  LinkState s;
  s.pos = v3(p[0], p[1], p[2]); s.rot = q4(p[3], p[4], p[5], p[6]); s.vel = v3(p[7], p[8], p[9]); s.ang = v3(p[10], p[11], p[12]);
  return s;
}
static void wr(float* p, const LinkState& s) {
  p[0] = s.pos.x; p[1] = s.pos.y; p[2] = s.pos.z; p[3] = s.rot.w; p[4] = s.rot.x; p[5] = s.rot.y; p[6] = s.rot.z;
  p[7] = s.vel.x; p[8] = s.vel.y; p[9] = s.vel.z; p[10] = s.ang.x; p[11] = s.ang.y; p[12] = s.ang.z;
}

// actuator torques of the (up to three) dofs of link l: gear * clip(action)
static void taus_of(const float* sys, int l, const float* act, float* tau) {
  const float* lt = link_tab(sys, l);
  const float* dt = dof_tab(sys, l);
  const int ai[3] = {(int)lt[L_ACT], (int)dt[D_ACT1], (int)dt[D_ACT2]};
  const float gear[3] = {lt[L_GEAR], dt[D_GEAR1], dt[D_GEAR2]};
  const int type = (int)lt[L_TYPE];
  const int nd = type == TYPE_FREE ? 0 : type_ndof(type);
  for (int d = 0; d < 3; ++d) {
    tau[d] = 0.0f;
    if (d < nd && ai[d] >= 0) tau[d] = gear[d] * fminf(fmaxf(act[ai[d]], lt[L_CTRL_LO]), lt[L_CTRL_HI]);
  }
}

static void obs_of(const float* sys, const LinkState* st, float* obs, float* root /*x,z,angle,ok,q1,qd1,qd2,site xyz*/,
                   const float* masses, const float* act) {
  const int L = (int)sys[H_N_LINKS], nq = (int)sys[H_N_Q], nqd = (int)sys[H_N_QD];
  float q[MAX_Q], qd[MAX_Q];
  for (int l = 0; l < L; ++l) {
    const float* lt = link_tab(sys, l);
    const int qi = (int)lt[L_QIDX], qdi = (int)lt[L_QDIDX], type = (int)lt[L_TYPE], parent = (int)lt[L_PARENT];
    if (type == TYPE_FREE) {
      const V3 o = link_origin(st[l], lt), vo = origin_velocity(st[l], lt), al = inv_rotate(st[l].ang, st[l].rot);
      q[qi] = o.x; q[qi + 1] = o.y; q[qi + 2] = o.z; q[qi + 3] = st[l].rot.w; q[qi + 4] = st[l].rot.x; q[qi + 5] = st[l].rot.y; q[qi + 6] = st[l].rot.z;
      qd[qdi] = vo.x; qd[qdi + 1] = vo.y; qd[qdi + 2] = vo.z; qd[qdi + 3] = al.x; qd[qdi + 4] = al.y; qd[qdi + 5] = al.z;
    } else {
      const JointOut jo = joint_resolve<true, true>(sys, lt, st[l], parent < 0, link_tab(sys, parent < 0 ? 0 : parent),
                                                    st[parent < 0 ? 0 : parent], 0.0f, 1.0f, dof_tab(sys, l));
      const int nd = type_ndof(type);
      for (int k = 0; k < nd; ++k) { q[qi + k] = jo.q[k]; qd[qdi + k] = jo.qd[k]; }
    }
  }
  const int ex = (int)sys[H_EXCLUDE_POS];
  const float clip = sys[H_QD_CLIP];
  int k = 0;
  const int kind = (int)sys[H_ENV];
  V3 site = v3(0, 0, 0);
  if (kind >= ENV_INVERTED_PENDULUM && kind <= ENV_REACHER) site = site_position(sys, st[(int)sys[H_SITE_LINK]], st[kind == ENV_REACHER ? 2 : 0]);
  float comx = 0.0f;
  if (kind == ENV_HUMANOID || kind == ENV_HUMANOIDSTANDUP) {
    // brax.envs.humanoid._get_obs, laid out as the kernel does: q[2:] | qd | cinert | cvel | actuator torques
    float rows[MAX_LINKS * LINK_WORDS], meff[MAX_LINKS];
    for (int l = 0; l < L; ++l) { wr(rows + l * LINK_WORDS, st[l]); meff[l] = eff_mass(masses[l], sys); }
    V3 com;
    const float msum = body_com(rows, meff, L, com);
    comx = com.x;
    for (int i = 2; i < nq; ++i) obs[k++] = q[i];
    for (int i = 0; i < nqd; ++i) obs[k++] = qd[i];
    for (int l = 0; l < L; ++l) { link_cinert(sys, link_tab(sys, l), st[l], meff[l], com, obs + k); k += 10; }
    for (int l = 0; l < L; ++l) { link_cvel(st[l], meff[l], msum, obs + k); k += 6; }
    for (int i = 0; i < nqd; ++i) obs[k + i] = 0.0f;
    for (int l = 0; l < L; ++l) {
      const float* lt = link_tab(sys, l);
      if ((int)lt[L_TYPE] == TYPE_FREE) continue;
      float tau[3];
      taus_of(sys, l, act, tau);
      for (int d = 0; d < type_ndof((int)lt[L_TYPE]); ++d) obs[k + (int)lt[L_QDIDX] + d] = tau[d];
    }
  } else if (kind == ENV_PUSHER) {
    float rows[MAX_LINKS * LINK_WORDS];
    for (int l = 0; l < L; ++l) wr(rows + l * LINK_WORDS, st[l]);
    for (int i = 0; i < 23; ++i) obs[i] = pusher_obs_entry(sys, i, q, qd, rows);
    site = pusher_distances(sys, rows);
  } else if (kind == ENV_INVERTED_DOUBLE_PENDULUM || kind == ENV_REACHER) {
    const int D = kind == ENV_REACHER ? 11 : 8;
    for (int i = 0; i < D; ++i) obs[i] = special_obs_entry(kind, i, q, qd, site);
  } else {
    for (int i = ex; i < nq; ++i) obs[k++] = q[i];
    for (int i = 0; i < nqd; ++i) { float v = qd[i]; if (clip > 0.0f) v = fminf(fmaxf(v, -clip), clip); obs[k++] = v; }
  }
  root[4] = q[1]; root[5] = qd[1]; root[6] = nqd > 2 ? qd[2] : 0.0f; root[7] = site.x; root[8] = site.y; root[9] = site.z;
  bool ok = true;
  for (int i = 2; i < nq; ++i) ok = ok && q[i] > -100.0f && q[i] < 100.0f;
  for (int i = 0; i < nqd; ++i) ok = ok && qd[i] > -100.0f && qd[i] < 100.0f;
  const V3 o0 = link_origin(st[0], link_tab(sys, 0));
  root[0] = kind == ENV_HUMANOID ? comx : o0.x; root[1] = o0.z; root[2] = ((int)link_tab(sys, 0)[L_TYPE] == TYPE_PLANAR) ? q[2] : 0.0f; root[3] = ok ? 1.0f : 0.0f;
}

static JointOut joint_resolve_shared_origins(const float* sys, const float* lt, const LinkState& c, bool world_parent,
                                             const float* plt, const LinkState& p, const float* tau, float stiffness_scale,
                                             const float* dt, V3 org_c, V3 org_p) {
  const V3 origins[2] = {org_c, org_p};
  return joint_resolve<true, true>(sys, lt, c, world_parent, plt, p, tau[0], stiffness_scale, parent_anchor(lt), joint_flags(lt),
                                   dt, tau[1], tau[2], origins);
}

// one env-step of all envs; FAST selects the reformulations of the FMA build (world-frame hinge, unit-inertia shortcut)
template <bool FAST>
static void step_all(const float* sys, int n, float* state, int words, const float* ctx, int n_ctx, const float* actions,
                     int* elapsed, int max_steps, int autoreset, const float* first_state, const float* first_obs, float* obs,
                     int D, float* reward, uint8_t* done_out, int stock_contact);

static int g_fast = 0;

extern "C" {

void hc_brax_set_fast(int fast) { g_fast = fast; }

void hc_brax_init(const float* sys, int n, const float* q_all, const float* qd_all, float* state, int words, float* obs, int D,
                  const float* ctx, int n_ctx) {
  const int L = (int)sys[H_N_LINKS], nq = (int)sys[H_N_Q], nqd = (int)sys[H_N_QD];
  for (int e = 0; e < n; ++e) {
    LinkState st[MAX_LINKS];
    for (int l = 0; l < L; ++l) {
      const float* lt = link_tab(sys, l);
      const int parent = (int)lt[L_PARENT];
      st[l] = forward_link<true>(sys, lt, q_all + (size_t)e * nq, qd_all + (size_t)e * nqd, parent < 0,
                                 link_tab(sys, parent < 0 ? 0 : parent), st[parent < 0 ? 0 : parent], dof_tab(sys, l));
    }
    float* rows = state + (size_t)e * words;
    memset(rows, 0, sizeof(float) * words);
    for (int l = 0; l < L; ++l) wr(rows + l * LINK_WORDS, st[l]);
    float root[10], zero_act[MAX_ACT] = {0};
    obs_of(sys, st, obs + (size_t)e * D, root, ctx + (size_t)e * n_ctx + C_MASS0, zero_act);
  }
}

void hc_brax_step(const float* sys, int n, float* state, int words, const float* ctx, int n_ctx, const float* actions,
                  int* elapsed, int max_steps, int autoreset, const float* first_state, const float* first_obs, float* obs,
                  int D, float* reward, uint8_t* done_out, int stock_contact) {
  if (g_fast)
    step_all<true>(sys, n, state, words, ctx, n_ctx, actions, elapsed, max_steps, autoreset, first_state, first_obs, obs, D, reward,
                   done_out, stock_contact);
  else
    step_all<false>(sys, n, state, words, ctx, n_ctx, actions, elapsed, max_steps, autoreset, first_state, first_obs, obs, D, reward,
                    done_out, stock_contact);
}

}  // extern "C"

template <bool FAST>
static void step_all(const float* sys, int n, float* state, int words, const float* ctx, int n_ctx, const float* actions,
                     int* elapsed, int max_steps, int autoreset, const float* first_state, const float* first_obs, float* obs,
                     int D, float* reward, uint8_t* done_out, int stock_contact) {
  const int L = (int)sys[H_N_LINKS], P = (int)sys[H_N_POINTS], A = (int)sys[H_N_ACT], NF = (int)sys[H_N_FRAMES];
  for (int e = 0; e < n; ++e) {
    float* rows = state + (size_t)e * words;
    const float* c = ctx + (size_t)e * n_ctx;
    const float* act = actions + (size_t)e * A;
    LinkState st[MAX_LINKS];
    for (int l = 0; l < L; ++l) st[l] = rd(rows + l * LINK_WORDS);
    float before[10], after[10], ob[MAX_OBS_LARGE];
    obs_of(sys, st, ob, before, c + C_MASS0, act);
    float act_sq = 0.0f;
    for (int a = 0; a < A; ++a) act_sq += act[a] * act[a];
    LinkConst lcs[MAX_LINKS];
    for (int l = 0; l < L; ++l) lcs[l] = make_link_const(sys, link_tab(sys, l), c[C_MASS0 + l], c[C_ANG_DAMPING]);
    float reach[MAX_POINTS];
    for (int p = 0; p < P; ++p) reach[p] = contact_reach(point_tab(sys, p), link_tab(sys, (int)point_tab(sys, p)[0]));
    for (int f = 0; f < NF; ++f) {
      Wrench w[MAX_LINKS], pw[MAX_LINKS];
      V3 org[MAX_LINKS];  // link-frame origins, evaluated once per link by its owner (as the kernel shares them)
      for (int l = 0; l < L; ++l) org[l] = link_origin(st[l], link_tab(sys, l));
      for (int l = 0; l < L; ++l) {
        w[l].f = v3(0, 0, 0); w[l].t = v3(0, 0, 0); pw[l] = w[l];
        const float* lt = link_tab(sys, l);
        if ((int)lt[L_TYPE] == TYPE_FREE) continue;
        const int parent = (int)lt[L_PARENT];
        float tau[3];
        taus_of(sys, l, act, tau);
        const float* plt = link_tab(sys, parent < 0 ? 0 : parent);
        const LinkState& pst = st[parent < 0 ? 0 : parent];
        const int jt = (int)lt[L_TYPE];
        const bool revolute = jt == TYPE_HINGE || jt == TYPE_PLANAR || jt == TYPE_HINGE2 || jt == TYPE_HINGE3;
        const bool special = is_special_env((int)sys[H_ENV]);  // kernels: MODE_SPECIAL
        const JointOut jo = (FAST && revolute && !special)
                                ? joint_resolve_world<true>(sys, lt, st[l], parent < 0, pst, tau[0], c[C_STIFFNESS_SCALE],
                                                            parent_anchor_from_com(lt, plt, parent < 0), joint_flags(lt),
                                                            dof_tab(sys, l), tau[1], tau[2])
                                : (FAST ? joint_resolve<true, true>(sys, lt, st[l], parent < 0, plt, pst, tau[0], c[C_STIFFNESS_SCALE],
                                                                    dof_tab(sys, l), tau[1], tau[2])
                                        : joint_resolve_shared_origins(sys, lt, st[l], parent < 0, plt, pst, tau, c[C_STIFFNESS_SCALE],
                                                                       dof_tab(sys, l), org[l], org[parent < 0 ? 0 : parent]));
        w[l] = jo.child; pw[l] = jo.parent;
      }
      LinkState nx[MAX_LINKS];
      for (int l = 0; l < L; ++l) {
        Wrench tot = w[l];
        for (int k = l + 1; k < L; ++k)
          if ((int)link_tab(sys, k)[L_PARENT] == l) { tot.f = tot.f + pw[k].f; tot.t = tot.t + pw[k].t; }
        nx[l] = st[l];
        integrate_xdd<FAST>(nx[l], tot, sys, link_tab(sys, l), lcs[l], c[C_GRAVITY]);
      }
      ContactOut co[MAX_POINTS];
      for (int p = 0; p < P; ++p) {
        const float* pt = point_tab(sys, p);
        const int l = (int)pt[0];
        const float fr = (c[C_FRICTION] < 0.0f || stock_contact) ? pt[5] : c[C_FRICTION];
        const float el = (c[C_ELASTICITY] < 0.0f || stock_contact) ? pt[6] : c[C_ELASTICITY];
        if (nx[l].pos.z > reach[p]) {  // the kernel's early-out: out of reach of the ground
          co[p].p = v3(0, 0, 0); co[p].t = v3(0, 0, 0); co[p].active = 0.0f;
          continue;
        }
        co[p] = contact_resolve<FAST>(sys, pt, link_tab(sys, l), nx[l], lcs[l], fr, el, org[l]);
      }
      // body-vs-body pairs (pusher): each writes the two candidate rows reserved for it
      for (int k = 0; k < (int)sys[OFF_PAIR + X_N_PAIRS]; ++k) {
        const float* pr = pair_tab(sys, k);
        const int la = (int)pr[R_LINK_A], lb = (int)pr[R_LINK_B], ra = (int)pr[R_ROW_A], rb = (int)pr[R_ROW_B];
        const float* pta = point_tab(sys, ra);
        const float fr = (c[C_FRICTION] < 0.0f || stock_contact) ? pta[5] : c[C_FRICTION];
        const float el = (c[C_ELASTICITY] < 0.0f || stock_contact) ? pta[6] : c[C_ELASTICITY];
        const PairOut po = pair_resolve(sys, pr, link_tab(sys, la), link_tab(sys, lb), nx[la], nx[lb], lcs[la], lcs[lb], fr, el);
        co[ra].p = po.p; co[ra].t = po.ta; co[ra].active = po.active;
        co[rb].p = v3(0, 0, 0) - po.p; co[rb].t = po.tb; co[rb].active = po.active;
      }
      for (int l = 0; l < L; ++l) {
        const float* lt = link_tab(sys, l);
        V3 ps = v3(0, 0, 0), ts = v3(0, 0, 0);
        float na = 0.0f;
        for (int k = (int)lt[L_FIRST_PT]; k < (int)lt[L_FIRST_PT] + (int)lt[L_N_PT]; ++k) { ps = ps + co[k].p; ts = ts + co[k].t; na += co[k].active; }
        integrate_xdv<FAST>(nx[l], ps, ts, na, lt, lcs[l]);
        integrate_pose(nx[l], sys[H_DT]);
        st[l] = nx[l];
      }
    }
    obs_of(sys, st, ob, after, c + C_MASS0, act);
    const float dt_env = sys[H_DT] * sys[H_N_FRAMES];
    const float xvel = (after[0] - before[0]) / dt_env;
    const int kind = (int)sys[H_ENV];
    bool healthy = true;
    if (kind == ENV_ANT || kind == ENV_HUMANOID) healthy = !(after[1] < sys[H_HEALTHY_Z_MIN]) && !(after[1] > sys[H_HEALTHY_Z_MAX]);
    else if (kind == ENV_HOPPER)
      healthy = after[3] > 0.5f && sys[H_HEALTHY_Z_MIN] < after[1] && after[1] < sys[H_HEALTHY_Z_MAX] &&
                sys[H_ANGLE_MIN] < after[2] && after[2] < sys[H_ANGLE_MAX];
    else if (kind == ENV_WALKER2D)
      healthy = !(after[1] < sys[H_HEALTHY_Z_MIN]) && !(after[1] > sys[H_HEALTHY_Z_MAX]) && !(after[2] > sys[H_ANGLE_MAX]) &&
                !(after[2] < sys[H_ANGLE_MIN]);
    float r = sys[H_FORWARD_WEIGHT] * xvel + sys[H_HEALTHY_REWARD] - sys[H_CTRL_COST] * act_sq;
    bool done = sys[H_TERMINATE] > 0.0f && !healthy;
    if (kind == ENV_PUSHER) {
      pusher_outcome(sys, v3(before[7], before[8], before[9]), act_sq, r, done);
    } else if (kind == ENV_HUMANOIDSTANDUP) {  // uph_cost + 1 - quad_ctrl_cost, never done
      r = (after[1] - 0.0f) / dt_env + sys[H_HEALTHY_REWARD] - sys[H_CTRL_COST] * act_sq;
      done = false;
    } else if (kind >= ENV_INVERTED_PENDULUM && kind <= ENV_REACHER) special_outcome(kind, after[4], after[5], after[6], v3(after[7], after[8], after[9]), act_sq, r, done);
    elapsed[e] += 1;
    if (max_steps > 0 && elapsed[e] >= max_steps) done = true;
    for (int l = 0; l < L; ++l) wr(rows + l * LINK_WORDS, st[l]);
    memcpy(obs + (size_t)e * D, ob, sizeof(float) * D);
    if (done && autoreset) {
      memcpy(rows, first_state + (size_t)e * words, sizeof(float) * words);
      memcpy(obs + (size_t)e * D, first_obs + (size_t)e * D, sizeof(float) * D);
      elapsed[e] = 0;
    }
    reward[e] = r;
    done_out[e] = done ? 1 : 0;
  }
}
