// Brax-locomotion kernels (sm_100a): ONE WARP PER 1, 2, 3 OR 4 ENV INSTANCES.
//
// Each env owns LPE = 32/E consecutive lanes of a warp (E = 4: Halfcheetah, Hopper, Walker2d, pendulums, reacher;
// E = 3: Ant, pusher; E = 2: humanoids). Sub-lanes 0..L-1 own the links of the body (13-float centre-of-mass state in
// registers) and, in ceil(P/LPE) passes, the ground-contact candidate points of the collision phase;
// the exchange between the lanes of an env (parent/child states, joint reactions, contact
// impulses, loop invariants) goes through a 2.7 KB shared-memory scratch guarded by __syncwarp.
// The system table (4.1 KB: link frames, inertias, joint limits, contact points, dof rows, contact pairs, tunables) is
// staged once per CTA into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier); each
// env's context scalars (gravity, friction, elasticity, ang_damping, stiffness scale, link masses)
// are staged with one coalesced load and then read as shared-memory broadcasts.
//
// This is FP32-issue / latency bound work (~6e4 flops per Ant env-step against ~1.1 kB of HBM
// traffic): no tensor cores. One launch = n_frames spring substeps (+ env layer); the fused
// rollout keeps the link state in registers across K env-steps.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "engine.h"
#include "physics_brax.h"

// Two builds of this file go into libcarlb.so (carl_b200/build.py): the STRICT variant (-fmad=false: every
// product and sum rounded separately, which is what reproduces the float32 restatement of the reference
// arithmetic to ~1e-6 per env-step) and, through brax_fma.cu, the FMA variant (ptxas contracts a*b+c into FFMA,
// as XLA does for the reference's own kernels: ~25 % fewer issued instructions, results at the float32 round-off
// floor of the algorithm, not bit-comparable with the strict oracle). Each variant lives in its own inner
// namespace so that its kernels are distinct symbols; carlb_brax_set_arithmetic selects per handle.
#ifdef CARLB_BRAX_FMA_BUILD
#define CARLB_BRAX_VARIANT fma_variant
#else
#define CARLB_BRAX_VARIANT strict_variant
#endif

namespace carlb {
namespace CARLB_BRAX_VARIANT {
using namespace brax;

// FAST arithmetic of the FMA build: the world-frame hinge and the unit-inertia shortcut (physics_brax.h) -- identical
// mathematics, different rounding; the strict build keeps the restated reference order everywhere.
#ifdef CARLB_BRAX_FMA_BUILD
constexpr bool FAST = true;
#else
constexpr bool FAST = false;
#endif

// Warps (= env instances) per CTA. 4 x 128-thread CTAs per SM is the default; for batches of several
// thousand envs a 7-warp CTA (2 resident per SM = 14 warps) makes the grid an almost exact multiple
// of the machine: 8 192 envs = 3.96 waves of 148 SMs x 14 warps instead of 3.46 waves of 16.

struct BraxSys {        // device-resident per-handle system table (+ static facts)
  float table[TABLE_FLOATS];
};

struct BraxSeg {
  int n;
  int max_steps;
  int autoreset;
  int state_words;   // padded 13*L
  int obs_dim;
  int n_ctx;         // 5 + L
  int act_dim;
  long long global_offset;
  uint64_t seed;
  int reset_rng;     // CARLB_RESET_PHILOX / CARLB_RESET_JAX
  uint32_t batch;    // JAX reset stream: size of the (global) batch the reference's VmapWrapper would split the key over
  const float* sys;  // BraxSys::table in global memory
  float* state;      // [n][state_words]
  const float* ctx;  // [n][n_ctx]  (AoS per env: one coalesced load + shuffle broadcast)
  int32_t* elapsed;
  uint64_t* episode; // [n] reset counter (rng buffer row 0)
  float* obs;        // [n][D]
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  float* final_obs;
  float* first_state;
  float* first_obs;
  GatherDev gth;     // fused cross-GPU obs gather of this launch (common.cuh; always an immediate push here)
  // zero-copy host mirrors (carlb_env_step_host with page-locked buffers): the single-step kernel also
  // stores obs / reward / flags straight into mapped host memory (posted PCIe writes behind the physics)
  float* host_obs;
  float* host_reward;
  uint8_t* host_terminated;
  uint8_t* host_truncated;
};

// Per-env shared-memory scratch (2.7 KB): the exchange medium between the lanes of one env.
struct EnvScratch {
  float ls[MAX_LINKS * LINK_WORDS];  // link states (also the coalesced I/O staging of the state row)
  float pw[MAX_LINKS * 6];           // reaction wrench of joint l on its parent
  float org[MAX_LINKS * 3];          // link-frame origins of this substep (joint phase -> contact phase)
  float co[MAX_POINTS * 7];          // contact outputs: impulse(3) angular impulse(3) active
  float lc[MAX_LINKS * 6];           // per-link loop invariants: inv_mass, inv_idiag(3), vel_decay, ang_decay
  float q[MAX_Q];
  float qd[MAX_Q];
  float act[16];
  float obs[64];
  float ctx[24];                     // this env's context scalars (gravity, friction, ..., link masses)
};
// The humanoids' scratch (3.6 KB): room for the 244-entry observation and 17 actions, plus the effective link
// masses that their observation and centre-of-mass reward read. Same leading fields as EnvScratch.
struct EnvScratchH {
  float ls[MAX_LINKS * LINK_WORDS];
  float pw[MAX_LINKS * 6];
  float org[MAX_LINKS * 3];
  float co[MAX_POINTS * 7];
  float lc[MAX_LINKS * 6];
  float q[MAX_Q];
  float qd[MAX_Q];
  float act[MAX_ACT];
  float obs[MAX_OBS_LARGE];
  float ctx[24];
  float meff[MAX_LINKS];             // mass^(1 - spring_mass_scale) per link
};
template <int M> struct ScratchOf { typedef EnvScratch type; };
template <> struct ScratchOf<MODE_HUMANOID> { typedef EnvScratchH type; };

// ---- TMA bulk copy + mbarrier (PTX) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Stage the system table into shared memory: one elected thread issues the bulk copy, everyone
// waits on the mbarrier phase.
__device__ __forceinline__ void stage_system(float* sys_s, const float* sys_g, uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, TABLE_FLOATS * sizeof(float));
    bulk_g2s(sys_s, sys_g, TABLE_FLOATS * sizeof(float), bar);
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ LinkState read_link(const float* ls, int l) {
  const float* p = ls + l * LINK_WORDS;
  LinkState s;
  s.pos = v3(p[0], p[1], p[2]);
  s.rot = q4(p[3], p[4], p[5], p[6]);
  s.vel = v3(p[7], p[8], p[9]);
  s.ang = v3(p[10], p[11], p[12]);
  return s;
}
__device__ __forceinline__ void write_link(float* ls, int l, const LinkState& s) {
  float* p = ls + l * LINK_WORDS;
  p[0] = s.pos.x; p[1] = s.pos.y; p[2] = s.pos.z;
  p[3] = s.rot.w; p[4] = s.rot.x; p[5] = s.rot.y; p[6] = s.rot.z;
  p[7] = s.vel.x; p[8] = s.vel.y; p[9] = s.vel.z;
  p[10] = s.ang.x; p[11] = s.ang.y; p[12] = s.ang.z;
}
__device__ __forceinline__ LinkConst read_lc(const float* lc, int l) {
  const float* p = lc + l * 6;
  LinkConst c;
  c.inv_mass = p[0]; c.inv_idiag = v3(p[1], p[2], p[3]); c.vel_decay = p[4]; c.ang_decay = p[5];
  return c;
}

// ---- lane mapping -----------------------------------------------------------------------------------
// E env instances share one warp: sub-env `sub` owns the LPE = 32/E consecutive lanes
// [sub*LPE, (sub+1)*LPE); within it sub-lane `sl` owns link `sl` (sl < L) and, in pass k of the
// collision phase, contact point k*LPE + sl. A warp of an Ant batch (9 links, 25 points) would use
// 9 of 32 lanes in the joint / integration phases with E = 1; E = 3 uses 27.
template <int E>
struct Lanes {
  static constexpr int LPE = 32 / E;
};

struct LaneCtx {
  int L, P, n_frames;
  int sub, sl;       // sub-env of this lane and the lane's index inside it
  bool active;       // this lane's sub-env maps onto a real env instance
  bool is_link;
  const float* lt;   // my link row
  const float* plt;  // parent's link row
  int parent, type, act;
  float gravity, stiffness_scale;
  bool stock_contact;
  LinkConst lc;      // loop invariants of my link
  V3 anchor_p;       // my joint's anchor in the parent link frame (table-only)
  int jflags;        // JointFlags of my joint (table-only)
  const float* dt;   // my link's dof row (stacked hinges: humanoid kernels only)
  unsigned children; // bit k set: link k is a child of my link (table-only)
  V3 anchor_pc;      // anchor_p relative to the parent's centre of mass (FAST hinge; table-only)
};

// n_frames spring substeps for the sub-envs of this warp. `s` is this lane's link.
// `reach`: per contact candidate, the COM height above which it cannot touch the ground (contact_reach; per CTA).
template <int E, int M>
__device__ __forceinline__ void pipeline_steps(const float* sys, const LaneCtx& c, typename ScratchOf<M>::type& w, LinkState& s,
                                               const float* reach, float tau, float tau1 = 0.0f, float tau2 = 0.0f) {
  constexpr bool SP = M == MODE_SPECIAL || M == MODE_PUSHER, HU = M == MODE_HUMANOID, PU = M == MODE_PUSHER;
  constexpr int LPE = Lanes<E>::LPE;
  const float dt = sys[H_DT];
  const int n_pass = (c.P + LPE - 1) / LPE;
  // Every link publishes its state and (strict arithmetic) its link-frame origin once per substep, right after the
  // pose integration: the joint phase of the next substep reads the parent's origin instead of re-deriving it, and the
  // contact phase reads the candidate's. The FAST arithmetic works from the centres of mass and needs no origins.
  if (c.is_link) {
    write_link(w.ls, c.sl, s);
    if (!FAST) {
      const V3 o0 = link_origin(s, c.lt);
      float* og = w.org + c.sl * 3;
      og[0] = o0.x; og[1] = o0.y; og[2] = o0.z;
    }
  }
  for (int f = 0; f < c.n_frames; ++f) {
    __syncwarp();
    // joints: sub-lane l resolves the joint between link l and its parent
    Wrench wr;
    wr.f = v3(0, 0, 0); wr.t = v3(0, 0, 0);
    if (c.is_link && c.type != TYPE_FREE) {
      const bool world_parent = c.parent < 0;
      const int pi = world_parent ? 0 : c.parent;
      const LinkState ps = read_link(w.ls, pi);
      JointOut jo;
      if (FAST && !SP) {
        // every joint of the locomotion bodies and the humanoids is revolute: one world-frame path for all lanes
        jo = joint_resolve_world<HU>(sys, c.lt, s, world_parent, ps, tau, c.stiffness_scale, c.anchor_pc, c.jflags, c.dt, tau1, tau2);
      } else if (FAST) {
        jo = joint_resolve<SP, HU>(sys, c.lt, s, world_parent, c.plt, ps, tau, c.stiffness_scale, c.anchor_p, c.jflags, c.dt,
                                   tau1, tau2);
      } else {
        const V3 origins[2] = {ld3(w.org + c.sl * 3), ld3(w.org + pi * 3)};
        jo = joint_resolve<SP, HU>(sys, c.lt, s, world_parent, c.plt, ps, tau, c.stiffness_scale, c.anchor_p, c.jflags, c.dt,
                                   tau1, tau2, origins);
      }
      wr = jo.child;
      float* pw = w.pw + c.sl * 6;
      pw[0] = jo.parent.f.x; pw[1] = jo.parent.f.y; pw[2] = jo.parent.f.z;
      pw[3] = jo.parent.t.x; pw[4] = jo.parent.t.y; pw[5] = jo.parent.t.z;
    }
    __syncwarp();
    if (c.is_link) {
      // add the reactions of my children (ascending link order: deterministic sums)
      for (unsigned m = c.children; m != 0u; m &= m - 1u) {
        const float* pw = w.pw + (__ffs((int)m) - 1) * 6;
        wr.f = wr.f + v3(pw[0], pw[1], pw[2]);
        wr.t = wr.t + v3(pw[3], pw[4], pw[5]);
      }
      integrate_xdd<FAST>(s, wr, sys, c.lt, c.lc, c.gravity);
      write_link(w.ls, c.sl, s);
    }
    __syncwarp();
    // ground contacts: in pass k sub-lane sl resolves candidate point k*LPE + sl against the plane
    for (int pass = 0; pass < n_pass; ++pass) {
      const int slot = pass * LPE + c.sl;
      if (c.active && c.sl < LPE && slot < c.P) {
        // contact schedule: candidates near the ground come first, so the later passes mostly return at the
        // penetration test for every lane of the warp (same impulses, stored per candidate as before)
        const int p = (int)point_tab(sys, slot)[P_SCHED];
        const float* pt = point_tab(sys, p);
        const int pl = (int)pt[0];
        float* o = w.co + p * 7;
        if (w.ls[pl * LINK_WORDS + 2] > reach[p]) {
          o[6] = 0.0f;  // the link's COM is higher than the candidate can reach down: inactive (only the flag is read)
        } else {
          const LinkState ps = read_link(w.ls, pl);
          const float fr = (c.stock_contact || w.ctx[C_FRICTION] < 0.0f) ? pt[5] : w.ctx[C_FRICTION];
          const float el = (c.stock_contact || w.ctx[C_ELASTICITY] < 0.0f) ? pt[6] : w.ctx[C_ELASTICITY];
          const ContactOut co = contact_resolve<FAST>(sys, pt, link_tab(sys, pl), ps, read_lc(w.lc, pl), fr, el,
                                                      FAST ? v3(0, 0, 0) : ld3(w.org + pl * 3));
          o[0] = co.p.x; o[1] = co.p.y; o[2] = co.p.z; o[3] = co.t.x; o[4] = co.t.y; o[5] = co.t.z; o[6] = co.active;
        }
      }
    }
    if constexpr (PU) {
      // body-vs-body pairs (pusher): sub-lane k resolves pair k and writes the two candidate rows reserved for it
      const int n_pairs = (int)sys[OFF_PAIR + X_N_PAIRS];
      if (n_pairs > 0) {  // block-uniform
        __syncwarp();  // the ground passes initialised those rows
        if (c.active && c.sl < n_pairs) {
          const float* pr = pair_tab(sys, c.sl);
          const int la = (int)pr[R_LINK_A], lb = (int)pr[R_LINK_B], row_a = (int)pr[R_ROW_A], row_b = (int)pr[R_ROW_B];
          const float* pta = point_tab(sys, row_a);
          const float fr = (c.stock_contact || w.ctx[C_FRICTION] < 0.0f) ? pta[5] : w.ctx[C_FRICTION];
          const float el = (c.stock_contact || w.ctx[C_ELASTICITY] < 0.0f) ? pta[6] : w.ctx[C_ELASTICITY];
          const PairOut po = pair_resolve(sys, pr, link_tab(sys, la), link_tab(sys, lb), read_link(w.ls, la), read_link(w.ls, lb),
                                          read_lc(w.lc, la), read_lc(w.lc, lb), fr, el);
          float* oa = w.co + row_a * 7;
          float* ob = w.co + row_b * 7;
          oa[0] = po.p.x; oa[1] = po.p.y; oa[2] = po.p.z; oa[3] = po.ta.x; oa[4] = po.ta.y; oa[5] = po.ta.z; oa[6] = po.active;
          ob[0] = 0.0f - po.p.x; ob[1] = 0.0f - po.p.y; ob[2] = 0.0f - po.p.z; ob[3] = po.tb.x; ob[4] = po.tb.y; ob[5] = po.tb.z;
          ob[6] = po.active;
        }
      }
    }
    __syncwarp();
    if (c.is_link) {
      V3 ps = v3(0, 0, 0), ts = v3(0, 0, 0);
      float na = 0.0f;
      const int p0 = (int)c.lt[L_FIRST_PT], np = (int)c.lt[L_N_PT];
      for (int k = p0; k < p0 + np; ++k) {
        const float* o = w.co + k * 7;
        const float on = o[6];
        if (on != 0.0f) {  // an inactive candidate holds exact zeros: skipping it leaves the sums bit-identical
          ps = ps + v3(o[0], o[1], o[2]);
          ts = ts + v3(o[3], o[4], o[5]);
          na += on;
        }
      }
      integrate_xdv<FAST>(s, ps, ts, na, c.lt, c.lc);
      integrate_pose(s, dt);
      write_link(w.ls, c.sl, s);
      if (!FAST) {
        const V3 o0 = link_origin(s, c.lt);
        float* og = w.org + c.sl * 3;
        og[0] = o0.x; og[1] = o0.y; og[2] = o0.z;
      }
    }
  }
  __syncwarp();
}

// kinematics.inverse + env observation: fills w.obs[0..D) and returns the root facts of the sub-env
struct RootFacts {
  float x, z, angle;
  bool state_ok;  // Hopper: all |q[2:]|, |qd| < 100
  // bodies without a locomotion root (inverted pendulums, reacher): what their env layer reads
  float q1, qd1, qd2;  // pole angle; hinge rates (unclipped)
  V3 site;             // pendulum tip (world) / reacher: fingertip - target
};

// M == MODE_SPECIAL: the body may have slide joints / an env-specific observation or outcome (inverted pendulums,
// reacher); M == MODE_HUMANOID: stacked hinges and the humanoids' observation (brax.envs.humanoid._get_obs)
template <int E, int M>
__device__ __forceinline__ RootFacts compute_obs(const float* sys, const LaneCtx& c, typename ScratchOf<M>::type& w,
                                                 const LinkState& s) {
  constexpr int LPE = Lanes<E>::LPE;
  constexpr bool SP = M == MODE_SPECIAL || M == MODE_PUSHER, HU = M == MODE_HUMANOID, PU = M == MODE_PUSHER;
  if (c.is_link) write_link(w.ls, c.sl, s);
  __syncwarp();
  const int nq = (int)sys[H_N_Q], nqd = (int)sys[H_N_QD];
  if (c.is_link) {
    const int qi = (int)c.lt[L_QIDX], qdi = (int)c.lt[L_QDIDX];
    if (c.type == TYPE_FREE) {
      const V3 o = link_origin(s, c.lt), vo = origin_velocity(s, c.lt), al = inv_rotate(s.ang, s.rot);
      w.q[qi + 0] = o.x; w.q[qi + 1] = o.y; w.q[qi + 2] = o.z;
      w.q[qi + 3] = s.rot.w; w.q[qi + 4] = s.rot.x; w.q[qi + 5] = s.rot.y; w.q[qi + 6] = s.rot.z;
      w.qd[qdi + 0] = vo.x; w.qd[qdi + 1] = vo.y; w.qd[qdi + 2] = vo.z;
      w.qd[qdi + 3] = al.x; w.qd[qdi + 4] = al.y; w.qd[qdi + 5] = al.z;
    } else {
      const bool world_parent = c.parent < 0;
      const LinkState ps = read_link(w.ls, world_parent ? 0 : c.parent);
      const JointOut jo = joint_resolve<SP, HU>(sys, c.lt, s, world_parent, c.plt, ps, 0.0f, 1.0f, c.dt);
      const int nd = (SP || HU) ? type_ndof(c.type) : (c.type == TYPE_PLANAR ? 3 : 1);
      for (int k = 0; k < nd; ++k) {
        w.q[qi + k] = jo.q[k];
        w.qd[qdi + k] = jo.qd[k];
      }
    }
  }
  __syncwarp();
  const int ex = (int)sys[H_EXCLUDE_POS];
  const float clip = sys[H_QD_CLIP];
  const int kind = (int)sys[H_ENV];
  RootFacts r;
  r.site = v3(0, 0, 0);
  if (PU) {
    r.site = pusher_distances(sys, w.ls);
  } else if (SP && kind >= ENV_INVERTED_PENDULUM) {  // block-uniform: the locomotion bodies never enter
    r.site = site_position(sys, read_link(w.ls, (int)sys[H_SITE_LINK]), read_link(w.ls, kind == ENV_REACHER ? 2 : 0));
  }
  float com_x = 0.0f;
  if constexpr (HU) {
    // q[2:] | qd | cinert (10 per link) | cvel (6 per link) | actuator torques in qd layout
    const int n0 = (nq - 2) + nqd;
    if (c.sl < LPE)
      for (int i = c.sl; i < n0; i += LPE) w.obs[i] = i < nq - 2 ? w.q[2 + i] : w.qd[i - (nq - 2)];
    V3 com;
    const float msum = body_com(w.ls, w.meff, c.L, com);  // every lane: the same sums in the same order
    com_x = com.x;
    if (c.is_link) {
      const float m = w.meff[c.sl];
      link_cinert(sys, c.lt, s, m, com, w.obs + n0 + 10 * c.sl);
      link_cvel(s, m, msum, w.obs + n0 + 10 * c.L + 6 * c.sl);
      float* qf = w.obs + n0 + 16 * c.L;
      const int qdi = (int)c.lt[L_QDIDX];
      if (c.type == TYPE_FREE) {
        for (int k = 0; k < 6; ++k) qf[qdi + k] = 0.0f;
      } else {
        const int a1 = (int)c.dt[D_ACT1], a2 = (int)c.dt[D_ACT2];
        const float lo = c.lt[L_CTRL_LO], hi = c.lt[L_CTRL_HI];
        qf[qdi] = c.act >= 0 ? c.lt[L_GEAR] * fminf(fmaxf(w.act[c.act], lo), hi) : 0.0f;
        if (c.type == TYPE_HINGE2 || c.type == TYPE_HINGE3) qf[qdi + 1] = a1 >= 0 ? c.dt[D_GEAR1] * fminf(fmaxf(w.act[a1], lo), hi) : 0.0f;
        if (c.type == TYPE_HINGE3) qf[qdi + 2] = a2 >= 0 ? c.dt[D_GEAR2] * fminf(fmaxf(w.act[a2], lo), hi) : 0.0f;
      }
    }
  } else if (PU) {
    if (c.sl < LPE)
      for (int i = c.sl; i < 23; i += LPE) w.obs[i] = pusher_obs_entry(sys, i, w.q, w.qd, w.ls);
  } else if (SP && (kind == ENV_INVERTED_DOUBLE_PENDULUM || kind == ENV_REACHER)) {
    const int D = kind == ENV_REACHER ? 11 : 8;
    if (c.sl < LPE)
      for (int i = c.sl; i < D; i += LPE) w.obs[i] = special_obs_entry(kind, i, w.q, w.qd, r.site);
  } else {
    const int D = (nq - ex) + nqd;
    if (c.sl < LPE) {
      for (int i = c.sl; i < D; i += LPE) {
        float v;
        if (i < nq - ex) {
          v = w.q[ex + i];
        } else {
          v = w.qd[i - (nq - ex)];
          if (clip > 0.0f) v = fminf(fmaxf(v, -clip), clip);
        }
        w.obs[i] = v;
      }
    }
  }
  __syncwarp();
  // Hopper healthy_state: every entry of state_vec = q[2:] ++ qd inside (-100, 100) (read by all lanes)
  bool ok = true;
  for (int i = 2; i < nq; ++i) ok = ok && (w.q[i] > -100.0f) && (w.q[i] < 100.0f);
  for (int i = 0; i < nqd; ++i) ok = ok && (w.qd[i] > -100.0f) && (w.qd[i] < 100.0f);
  r.state_ok = ok;
  r.q1 = w.q[1];
  r.qd1 = w.qd[1];
  r.qd2 = w.qd[2];
  const float* lt0 = link_tab(sys, 0);
  const LinkState s0 = read_link(w.ls, 0);
  const V3 o0 = link_origin(s0, lt0);
  r.x = o0.x;
  if (HU && kind == ENV_HUMANOID) r.x = com_x;  // brax.envs.humanoid: velocity of the body's centre of mass
  r.z = o0.z;
  r.angle = ((int)lt0[L_TYPE] == TYPE_PLANAR) ? w.q[2] : 0.0f;
  return r;
}

// Env layer of brax.envs.{ant,half_cheetah,hopper}.step after the pipeline advanced.
template <int M>
__device__ __forceinline__ void env_outcome(const float* sys, const RootFacts& before, const RootFacts& after,
                                            float act_sq_sum, float& reward, bool& done) {
  const float dt_env = sys[H_DT] * sys[H_N_FRAMES];
  const float x_velocity = (after.x - before.x) / dt_env;
  const float forward_reward = sys[H_FORWARD_WEIGHT] * x_velocity;
  const int kind = (int)sys[H_ENV];
  constexpr bool SP = M == MODE_SPECIAL, HU = M == MODE_HUMANOID, PU = M == MODE_PUSHER;
  bool healthy = true;
  if (kind == ENV_ANT || (HU && kind == ENV_HUMANOID)) {
    healthy = !(after.z < sys[H_HEALTHY_Z_MIN]) && !(after.z > sys[H_HEALTHY_Z_MAX]);
  } else if (kind == ENV_HOPPER) {
    const bool hz = (sys[H_HEALTHY_Z_MIN] < after.z) && (after.z < sys[H_HEALTHY_Z_MAX]);
    const bool ha = (sys[H_ANGLE_MIN] < after.angle) && (after.angle < sys[H_ANGLE_MAX]);
    healthy = after.state_ok && hz && ha;
  } else if (kind == ENV_WALKER2D) {
    healthy = !(after.z < sys[H_HEALTHY_Z_MIN]) && !(after.z > sys[H_HEALTHY_Z_MAX]) &&
              !(after.angle > sys[H_ANGLE_MAX]) && !(after.angle < sys[H_ANGLE_MIN]);
  }
  const float ctrl_cost = sys[H_CTRL_COST] * act_sq_sum;
  reward = forward_reward + sys[H_HEALTHY_REWARD] - ctrl_cost;
  done = (sys[H_TERMINATE] > 0.0f) && !healthy;
  if (PU) pusher_outcome(sys, before.site, act_sq_sum, reward, done);
  else if (SP && kind >= ENV_INVERTED_PENDULUM) special_outcome(kind, after.q1, after.qd1, after.qd2, after.site, act_sq_sum, reward, done);
  if (HU && kind == ENV_HUMANOIDSTANDUP) {  // brax.envs.humanoidstandup.step: uph_cost + 1 - quad_ctrl_cost, never done
    reward = (after.z - 0.0f) / dt_env + sys[H_HEALTHY_REWARD] - ctrl_cost;
    done = false;
  }
}

template <int W, int E, int M = MODE_LOCO>
struct SmemLayoutT {
  float sys[TABLE_FLOATS];
  typename ScratchOf<M>::type env[W * E];
  uint64_t bar;
  float reach[MAX_POINTS];  // contact_reach of every candidate (table-only)
};

// Per-lane setup shared by the step and reset kernels: lane mapping, context staging (one strided
// coalesced load of the env's context row into its scratch, then shared-memory broadcast reads),
// loop invariants.
template <int E, class SC>
__device__ __forceinline__ LaneCtx make_lane_ctx(const float* sys, const BraxSeg& seg, SC& w, int env, bool active,
                                                 int sub, int sl, bool stock_contact) {
  constexpr int LPE = Lanes<E>::LPE;
  LaneCtx c;
  c.L = (int)sys[H_N_LINKS];
  c.P = (int)sys[H_N_POINTS];
  c.n_frames = (int)sys[H_N_FRAMES];
  c.sub = sub;
  c.sl = sl;
  c.active = active;
  c.is_link = active && sl < LPE && sl < c.L;
  const int l = c.is_link ? sl : 0;
  c.lt = link_tab(sys, l);
  c.parent = (int)c.lt[L_PARENT];
  c.plt = link_tab(sys, c.parent >= 0 ? c.parent : 0);
  c.type = (int)c.lt[L_TYPE];
  c.act = (int)c.lt[L_ACT];
  c.stock_contact = stock_contact;
  c.anchor_p = parent_anchor(c.lt);
  c.jflags = joint_flags(c.lt);
  c.anchor_pc = parent_anchor_from_com(c.lt, c.plt, c.parent < 0);
  c.dt = dof_tab(sys, l);
  c.children = 0u;
  if (c.is_link)
    for (int k = l + 1; k < c.L; ++k)
      if ((int)link_tab(sys, k)[L_PARENT] == l) c.children |= 1u << k;
  if (active && sl < LPE) {
    const float* row = seg.ctx + (size_t)env * seg.n_ctx;
    for (int i = sl; i < seg.n_ctx; i += LPE) w.ctx[i] = row[i];
  }
  __syncwarp();
  c.gravity = w.ctx[C_GRAVITY];
  c.stiffness_scale = w.ctx[C_STIFFNESS_SCALE];
  c.lc = make_link_const(sys, c.lt, active ? w.ctx[C_MASS0 + l] : 1.0f, active ? w.ctx[C_ANG_DAMPING] : 0.0f);
  if (c.is_link) {
    float* p = w.lc + sl * 6;
    p[0] = c.lc.inv_mass; p[1] = c.lc.inv_idiag.x; p[2] = c.lc.inv_idiag.y; p[3] = c.lc.inv_idiag.z;
    p[4] = c.lc.vel_decay; p[5] = c.lc.ang_decay;
    if constexpr (std::is_same<SC, EnvScratchH>::value) w.meff[sl] = eff_mass(w.ctx[C_MASS0 + l], sys);
  }
  if constexpr (std::is_same<SC, EnvScratchH>::value) {
    if (active && sl < LPE)
      for (int i = sl; i < MAX_ACT; i += LPE) w.act[i] = 0.0f;  // the observation reads the actuator torques: none yet
  }
  __syncwarp();
  return c;
}

// ------------------------------------------------------------------------------- step
// One env-step per sub-env: [AutoReset zeroing] -> n_frames substeps -> obs/reward/done ->
// EpisodeWrapper truncation -> AutoReset (state/obs replaced by the stored first ones where done).
// Every __syncwarp() is reached by all 32 lanes: per-env conditions only predicate memory traffic.
template <int W, int E, int M>
__device__ __forceinline__ void brax_step_body(const BraxSeg& seg, SmemLayoutT<W, E, M>& sm, const float* actions, int n_steps,
                                               uint64_t policy_seed, uint32_t step_base, const carlb_traj_t& traj,
                                               int stock_contact, unsigned int gseq) {
  constexpr int LPE = Lanes<E>::LPE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = min(lane / LPE, E - 1), sl = lane - sub * LPE;  // lanes beyond E*LPE idle in the last sub-env
  const bool lane_ok = lane < E * LPE;
  const int env = (blockIdx.x * W + warp) * E + sub;
  const bool active = lane_ok && env < seg.n;
  const float* sys = sm.sys;
  constexpr bool HU = M == MODE_HUMANOID;
  typename ScratchOf<M>::type& w = sm.env[warp * E + sub];
  const LaneCtx c = make_lane_ctx<E>(sys, seg, w, env, active, sub, lane_ok ? sl : LPE, stock_contact != 0);
  const int words = seg.state_words, D = seg.obs_dim, A = seg.act_dim;
  float* state_row = seg.state + (size_t)(active ? env : 0) * words;
  if (active && c.sl < LPE)
    for (int i = c.sl; i < words; i += LPE) w.ls[i] = state_row[i];
  __syncwarp();
  LinkState s = read_link(w.ls, c.is_link ? c.sl : 0);
  __syncwarp();
  int el = active ? seg.elapsed[env] : 0;
  const uint64_t gid = (uint64_t)(seg.global_offset + env);
  float reward = 0.0f;
  bool done = false;
  RootFacts before = compute_obs<E, M>(sys, c, w, s);
  for (int t = 0; t < n_steps; ++t) {
    // actions of this step: given tensor [n][A] (single step) / [K][n][A] (rollout) or Philox policy
    float a = 0.0f;
    if constexpr (HU) {
      // 17 actions for 16 (or 32) lanes: a strided loop; the trajectory copy is stored as they are read
      if (active && c.sl < LPE) {
        for (int k = c.sl; k < A; k += LPE) {
          float ak;
          if (actions != nullptr) {
            ak = actions[((size_t)t * seg.n + env) * A + k];
          } else {
            const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), step_base + (uint32_t)t,
                                            0x42524158u + (uint32_t)k, (uint32_t)policy_seed, (uint32_t)(policy_seed >> 32));
            ak = sys[H_ACT_SCALE] * (2.0f * u32_to_unit_float(r.v[0]) - 1.0f);
          }
          w.act[k] = ak;
          if (traj.actions != nullptr && (n_steps > 1 || traj.obs != nullptr))
            static_cast<float*>(traj.actions)[((size_t)t * seg.n + env) * A + k] = ak;
        }
      }
    } else if (active && c.sl < A) {
      if (actions != nullptr) {
        a = actions[((size_t)t * seg.n + env) * A + c.sl];
      } else {
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), step_base + (uint32_t)t,
                                        0x42524158u + (uint32_t)c.sl, (uint32_t)policy_seed, (uint32_t)(policy_seed >> 32));
        a = sys[H_ACT_SCALE] * (2.0f * u32_to_unit_float(r.v[0]) - 1.0f);  // uniform over the action space
      }
      w.act[c.sl] = a;
    }
    __syncwarp();
    float act_sq = 0.0f;
    for (int k = 0; k < A; ++k) act_sq += w.act[k] * w.act[k];  // jp.sum(jp.square(action)), in order
    float tau = 0.0f, tau1 = 0.0f, tau2 = 0.0f;
    if (c.is_link && c.act >= 0) tau = c.lt[L_GEAR] * fminf(fmaxf(w.act[c.act], c.lt[L_CTRL_LO]), c.lt[L_CTRL_HI]);
    if constexpr (HU) {  // dofs 1 and 2 of a stacked hinge
      const int a1 = (int)c.dt[D_ACT1], a2 = (int)c.dt[D_ACT2];
      if (c.is_link && a1 >= 0) tau1 = c.dt[D_GEAR1] * fminf(fmaxf(w.act[a1], c.lt[L_CTRL_LO]), c.lt[L_CTRL_HI]);
      if (c.is_link && a2 >= 0) tau2 = c.dt[D_GEAR2] * fminf(fmaxf(w.act[a2], c.lt[L_CTRL_LO]), c.lt[L_CTRL_HI]);
    }
    pipeline_steps<E, M>(sys, c, w, s, sm.reach, tau, tau1, tau2);
    const RootFacts after = compute_obs<E, M>(sys, c, w, s);
    env_outcome<M>(sys, before, after, act_sq, reward, done);
    // EpisodeWrapper: steps += 1; done = where(steps >= episode_length, 1, done)
    el += 1;
    if (seg.max_steps > 0 && el >= seg.max_steps) done = true;
    before = after;
    const bool do_reset = active && done && seg.autoreset != CARLB_AUTORESET_NONE;
    if (__any_sync(0xffffffffu, do_reset)) {  // warp-uniform: the barriers inside are reached by all lanes
      if (do_reset && seg.final_obs != nullptr && n_steps == 1 && c.sl < LPE)
        for (int i = c.sl; i < D; i += LPE) seg.final_obs[(size_t)env * D + i] = w.obs[i];
      __syncwarp();
      // AutoResetWrapper: pipeline_state and obs <- the ones stored at reset (predicated per sub-env)
      if (do_reset && c.sl < LPE) {
        const float* fs = seg.first_state + (size_t)env * words;
        for (int i = c.sl; i < words; i += LPE) w.ls[i] = fs[i];
      }
      __syncwarp();
      if (do_reset) s = read_link(w.ls, c.is_link ? c.sl : 0);
      __syncwarp();
      // root facts of the restored state (also recomputes the unchanged obs of the other sub-envs)
      const RootFacts fresh = compute_obs<E, M>(sys, c, w, s);
      if (do_reset) {
        before = fresh;
        el = 0;
      }
      if (do_reset && c.sl < LPE)
        for (int i = c.sl; i < D; i += LPE) w.obs[i] = seg.first_obs[(size_t)env * D + i];
      __syncwarp();
    }
    if (active && (n_steps > 1 || traj.obs != nullptr)) {
      const size_t row = (size_t)t * seg.n + env;
      if (traj.obs != nullptr && c.sl < LPE)
        for (int i = c.sl; i < D; i += LPE) traj.obs[row * D + i] = w.obs[i];
      if (!HU && traj.actions != nullptr && c.sl < A) static_cast<float*>(traj.actions)[row * A + c.sl] = a;
      if (c.sl == 0) {
        if (traj.reward != nullptr) traj.reward[row] = reward;
        if (traj.done != nullptr) traj.done[row] = done ? 1 : 0;
      }
    }
    __syncwarp();
  }
  // write back: state row (coalesced through the scratch), obs, scalars
  if (c.is_link) write_link(w.ls, c.sl, s);
  __syncwarp();
  if (active && c.sl < LPE) {
    for (int i = c.sl; i < words; i += LPE) state_row[i] = w.ls[i];
    for (int i = c.sl; i < D; i += LPE) {
      const float v = w.obs[i];
      seg.obs[(size_t)env * D + i] = v;
      if (seg.gth.n_peers > 0) gather_store_elem(seg.gth, gseq, (size_t)(seg.global_offset + env) * D + i, v);
    }
    if (c.sl == 0) {
      seg.reward[env] = reward;
      seg.terminated[env] = done ? 1 : 0;  // CARL maps brax `done` (incl. the time limit) to terminated
      seg.truncated[env] = 0;              // and truncated = False always (wrappers.py:75-78)
      seg.elapsed[env] = el;
      if (seg.host_obs != nullptr) {
        seg.host_reward[env] = reward;
        seg.host_terminated[env] = done ? 1 : 0;
        seg.host_truncated[env] = 0;
      }
    }
  }
  if (seg.host_obs != nullptr) {
    // zero-copy obs mirror: the E envs of this warp are consecutive, so their E * D floats are ONE contiguous run in
    // the host array -- the 32 lanes store it 128 B at a time (PCIe likes full-size posted writes; per-env sub-lane
    // stores would go out as 40-byte fragments)
    __syncwarp();
    const int env0 = (blockIdx.x * W + warp) * E;
    const int n_run = min(E, seg.n - env0) * D;
    float* dst = seg.host_obs + (size_t)env0 * D;
    for (int j = lane; j < n_run; j += 32) dst[j] = sm.env[warp * E + j / D].obs[j % D];
  }
}

template <int W, int E, int M>
__global__ void __launch_bounds__(W * 32, W == 4 ? (E >= 3 ? 5 : 4) : 2) brax_step_kernel(const __grid_constant__ BraxSeg seg,
                                                                          const float* actions, int n_steps,
                                                                          uint64_t policy_seed, uint32_t step_base,
                                                                          const carlb_traj_t traj, int stock_contact) {
  typedef SmemLayoutT<W, E, M> SmemLayout;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  stage_system(sm.sys, seg.sys, &sm.bar);
  if (threadIdx.x < (int)sm.sys[H_N_POINTS]) {
    const float* pt = point_tab(sm.sys, threadIdx.x);
    sm.reach[threadIdx.x] = contact_reach(pt, link_tab(sm.sys, (int)pt[0]));
  }
  __syncthreads();
  const unsigned int gseq = gather_begin(seg.gth);
  const int warp = threadIdx.x >> 5;
  if ((blockIdx.x * W + warp) * E < seg.n)  // warp-uniform: at least one sub-env of this warp is real
    brax_step_body<W, E, M>(seg, sm, actions, n_steps, policy_seed, step_base, traj, stock_contact, gseq);
  gather_epilogue_immediate(seg.gth, gseq);
}

// ------------------------------------------------------------------------------ reset
// Env.reset: q = init_q + U(+-noise), qd = noise * N(0,1) (Hopper: both uniform), forward
// kinematics (pipeline_init), obs; stores the first state/obs for AutoResetWrapper. One env per warp.
template <int M>
__global__ void __launch_bounds__(128) brax_reset_kernel(const __grid_constant__ BraxSeg seg, const uint8_t* mask,
                                                         const float* q_in, const float* qd_in) {
  typedef SmemLayoutT<4, 1, M> SmemLayout;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  stage_system(sm.sys, seg.sys, &sm.bar);
  const unsigned int gseq = gather_begin(seg.gth);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * 4 + warp;
  if (env < seg.n) {
    if (mask != nullptr && mask[env] == 0) {
      // not reset: with a fused gather attached the current row still has to reach the new slot
      if (seg.gth.n_peers > 0)
        for (int i = lane; i < seg.obs_dim; i += 32)
          gather_store_elem(seg.gth, gseq, (size_t)(seg.global_offset + env) * seg.obs_dim + i, seg.obs[(size_t)env * seg.obs_dim + i]);
    } else {
      const float* sys = sm.sys;
      typename ScratchOf<M>::type& w = sm.env[warp];
      const LaneCtx c = make_lane_ctx<1>(sys, seg, w, env, true, 0, lane, true);
      const int nq = (int)sys[H_N_Q], nqd = (int)sys[H_N_QD];
      const uint64_t gid = (uint64_t)(seg.global_offset + env);
      const uint32_t episode = (uint32_t)seg.episode[env];
      const float noise = sys[H_RESET_NOISE], qd_noise = sys[H_QD_NOISE];
      const bool hopper = sys[H_QD_UNIFORM] > 0.0f;  // Hopper / Walker2d / pendulum / reacher: qd ~ U(+-noise), else noise * N(0,1)
      const bool reacher = (int)sys[H_ENV] == ENV_REACHER;
      const bool pusher = (int)sys[H_ENV] == ENV_PUSHER;
      if (seg.reset_rng == CARLB_RESET_JAX && (q_in == nullptr || qd_in == nullptr)) {
        // the reference's own stream: `rng, rng1, rng2 = jax.random.split(rng, 3)` on the key this env receives at
        // its `episode`-th reset, q = init_q + uniform(rng1, (nq,), -noise, noise), qd = noise * normal(rng2, (nqd,))
        // (Ant, Halfcheetah, InvertedDoublePendulum) or uniform(rng2, (nqd,), -noise, noise) (Hopper, Walker2d,
        // InvertedPendulum, Reacher) -- brax 0.12.1 envs/*.py `reset`
        const JaxKey ek = jax_env_reset_key(seg.seed, episode, seg.batch, (uint32_t)gid);
        const JaxKey rng0 = jax_split(ek, 3u, 0u), rng1 = jax_split(ek, 3u, 1u), rng2 = jax_split(ek, 3u, 2u);
        if (lane < nq && q_in == nullptr)
          w.q[lane] = sys[OFF_INIT_Q + lane] + jax_uniform(rng1, (uint32_t)nq, (uint32_t)lane, -noise, noise);
        if (lane < nqd && qd_in == nullptr)
          w.qd[lane] = hopper ? jax_uniform(rng2, (uint32_t)nqd, (uint32_t)lane, -qd_noise, qd_noise)
                              : qd_noise * jax_normal(rng2, (uint32_t)nqd, (uint32_t)lane);
        if (reacher && q_in == nullptr && lane >= 2 && lane < 4) {
          // brax.envs.reacher._random_target(rng): rng, rng1, rng2 = split(rng, 3); dist = 0.2 U(rng1); ang = 2 pi U(rng2)
          const JaxKey t1 = jax_split(rng0, 3u, 1u), t2 = jax_split(rng0, 3u, 2u);
          const float dist = 0.2f * jax_uniform(t1, 1u, 0u, 0.0f, 1.0f);
          const float ang = 6.283185307179586f * jax_uniform(t2, 1u, 0u, 0.0f, 1.0f);
          w.q[lane] = lane == 2 ? dist * cosf(ang) : dist * sinf(ang);
          w.qd[lane] = 0.0f;
        }
        if (pusher && q_in == nullptr && lane >= 7 && lane < 11) {
          // brax.envs.pusher.reset: the object starts at (U(rng, -0.3, -1e-6), U(rng1, -0.2, 0.2)) on its two slides, pushed
          // out to 0.17 from the goal if closer; the goal offsets and the last four rates are zero
          const float c0 = jax_uniform(rng0, 1u, 0u, -0.3f, -1e-6f), c1 = jax_uniform(rng1, 1u, 0u, -0.2f, 0.2f);
          const float nrm = sqrtf(c0 * c0 + c1 * c1);
          const float scale = nrm < 0.17f ? 0.17f / nrm : 1.0f;
          w.q[lane] = lane == 7 ? c0 * scale : (lane == 8 ? c1 * scale : 0.0f);
          w.qd[lane] = 0.0f;
        }
        if (lane < nq && q_in != nullptr) w.q[lane] = q_in[(size_t)env * nq + lane];
        if (lane < nqd && qd_in != nullptr) w.qd[lane] = qd_in[(size_t)env * nqd + lane];
      } else {
        if (lane < nq) {
          w.q[lane] = (q_in != nullptr) ? q_in[(size_t)env * nq + lane]
                                        : sys[OFF_INIT_Q + lane] + reset_uniform(seg.seed, gid, episode, (uint32_t)lane, -noise, noise);
        }
        if (lane < nqd) {
          w.qd[lane] = (qd_in != nullptr) ? qd_in[(size_t)env * nqd + lane]
                       : hopper           ? reset_uniform(seg.seed, gid, episode, 64u + (uint32_t)lane, -qd_noise, qd_noise)
                                          : qd_noise * reset_normal(seg.seed, gid, episode, (uint32_t)lane);
        }
        if (reacher && q_in == nullptr && lane >= 2 && lane < 4) {
          // brax.envs.reacher._random_target: dist = 0.2 U, ang = 2 pi U; q[2:] = target, qd[2:] = 0
          const float dist = 0.2f * reset_uniform(seg.seed, gid, episode, 128u, 0.0f, 1.0f);
          const float ang = 6.283185307179586f * reset_uniform(seg.seed, gid, episode, 129u, 0.0f, 1.0f);
          w.q[lane] = lane == 2 ? dist * cosf(ang) : dist * sinf(ang);
          w.qd[lane] = 0.0f;
        }
        if (pusher && q_in == nullptr && lane >= 7 && lane < 11) {
          const float c0 = reset_uniform(seg.seed, gid, episode, 128u, -0.3f, -1e-6f);
          const float c1 = reset_uniform(seg.seed, gid, episode, 129u, -0.2f, 0.2f);
          const float nrm = sqrtf(c0 * c0 + c1 * c1);
          const float scale = nrm < 0.17f ? 0.17f / nrm : 1.0f;
          w.q[lane] = lane == 7 ? c0 * scale : (lane == 8 ? c1 * scale : 0.0f);
          w.qd[lane] = 0.0f;
        }
      }
      __syncwarp();
      // forward kinematics down the tree (parents have smaller indices)
      LinkState s;
      s.pos = v3(0, 0, 0); s.rot = q4(1, 0, 0, 0); s.vel = v3(0, 0, 0); s.ang = v3(0, 0, 0);
      for (int l = 0; l < c.L; ++l) {
        if (lane == l) {
          const bool world_parent = c.parent < 0;
          const LinkState ps = read_link(w.ls, world_parent ? 0 : c.parent);
          s = forward_link<M == MODE_HUMANOID>(sys, c.lt, w.q, w.qd, world_parent, c.plt, ps, c.dt);
          write_link(w.ls, lane, s);
        }
        __syncwarp();
      }
      compute_obs<1, M>(sys, c, w, s);
      const int D = seg.obs_dim;
      for (int i = c.L * LINK_WORDS + lane; i < seg.state_words; i += 32) w.ls[i] = 0.0f;  // padding words
      __syncwarp();
      for (int i = lane; i < seg.state_words; i += 32) {
        seg.state[(size_t)env * seg.state_words + i] = w.ls[i];
        seg.first_state[(size_t)env * seg.state_words + i] = w.ls[i];
      }
      for (int i = lane; i < D; i += 32) {
        const float v = w.obs[i];
        seg.obs[(size_t)env * D + i] = v;
        seg.first_obs[(size_t)env * D + i] = v;
        if (seg.gth.n_peers > 0) gather_store_elem(seg.gth, gseq, (size_t)(seg.global_offset + env) * D + i, v);
      }
      if (lane == 0) {
        seg.reward[env] = 0.0f;
        seg.terminated[env] = 0;
        seg.truncated[env] = 0;
        seg.elapsed[env] = 0;
        seg.episode[env] = (uint64_t)episode + 1ull;
      }
    }
  }
  gather_epilogue_immediate(seg.gth, gseq);
}

// ---------------------------------------------------------------------------- host side
struct BraxHandle {
  int reset_rng = CARLB_RESET_PHILOX;
  uint32_t batch = 1;
  BraxSys* dev_sys = nullptr;
  float host_table[TABLE_FLOATS] = {};
  bool have_table = false;
  uint64_t seed = 0;
  int stock_contact = 0;
};

static void static_facts(int kind, int& L, int& nq, int& nqd, int& A) {
  switch (kind) {
    case KIND_BRAX_ANT: L = 9; nq = 15; nqd = 14; A = 8; break;
    case KIND_BRAX_HALFCHEETAH: L = 7; nq = 9; nqd = 9; A = 6; break;
    case KIND_BRAX_WALKER2D: L = 7; nq = 9; nqd = 9; A = 6; break;
    case KIND_BRAX_INVERTED_PENDULUM: L = 2; nq = 2; nqd = 2; A = 1; break;
    case KIND_BRAX_INVERTED_DOUBLE_PENDULUM: L = 3; nq = 3; nqd = 3; A = 1; break;
    case KIND_BRAX_REACHER: L = 3; nq = 4; nqd = 4; A = 2; break;
    case KIND_BRAX_HUMANOID:
    case KIND_BRAX_HUMANOIDSTANDUP: L = 11; nq = 24; nqd = 23; A = 17; break;
    case KIND_BRAX_PUSHER: L = 9; nq = 11; nqd = 11; A = 7; break;
    default: L = 4; nq = 6; nqd = 6; A = 3; break;
  }
}

int brax_query(int kind, carlb_env_info_t* o) {
  int L, nq, nqd, A;
  static_facts(kind, L, nq, nqd, A);
  const bool humanoid = kind == KIND_BRAX_HUMANOID || kind == KIND_BRAX_HUMANOIDSTANDUP;
  const int ex = (kind == KIND_BRAX_ANT || humanoid) ? 2 : ((kind == KIND_BRAX_INVERTED_PENDULUM || kind == KIND_BRAX_PUSHER) ? 0 : 1);
  o->kind = kind;
  o->state_words = ((LINK_WORDS * L + 3) / 4) * 4;
  o->obs_dim = (nq - ex) + nqd;
  if (kind == KIND_BRAX_INVERTED_DOUBLE_PENDULUM) o->obs_dim = 8;  // q0, sin, cos, clipped qd
  if (kind == KIND_BRAX_REACHER) o->obs_dim = 11;                  // cos, sin, target, qd[:2], tip - target
  if (humanoid) o->obs_dim = humanoid_obs_dim(nq, nqd, L);         // 244: q[2:], qd, cinert, cvel, actuator torques
  if (kind == KIND_BRAX_PUSHER) o->obs_dim = 23;                   // q[:7], qd[:7], three centres of mass
  o->act_dim = A;
  o->act_discrete = 0;
  o->n_actions = 0;
  o->n_param_rows = 5 + L;
  o->n_step_rows = 5 + L;
  o->default_max_steps = 1000;  // brax.envs.create(episode_length=1000)
  o->gym_reset_draws = 0;
  o->act_low = kind == KIND_BRAX_INVERTED_PENDULUM ? -3.0f : (humanoid ? -0.4f : (kind == KIND_BRAX_PUSHER ? -2.0f : -1.0f));  // BraxGymWrapper: Box(ctrl_range) (wrappers.py:48-50)
  o->act_high = -o->act_low;
  return CARLB_OK;
}

int brax_create(carlb_env* env) {
  BraxHandle* h = new BraxHandle();
  cudaError_t e = cudaMalloc(&h->dev_sys, sizeof(BraxSys));
  if (e != cudaSuccess) {
    delete h;
    set_error("cudaMalloc of the Brax system table failed: %s", cudaGetErrorString(e));
    return CARLB_ERR_CUDA;
  }
  env->brax_sys = h;
  return CARLB_OK;
}

void brax_destroy(carlb_env* env) {
  BraxHandle* h = static_cast<BraxHandle*>(env->brax_sys);
  if (h == nullptr) return;
  if (h->dev_sys) cudaFree(h->dev_sys);
  delete h;
  env->brax_sys = nullptr;
}

int brax_set_system(carlb_env* env, const float* table, int n_floats, int stock_contact) {
  BraxHandle* h = static_cast<BraxHandle*>(env->brax_sys);
  if (n_floats != TABLE_FLOATS) {
    set_error("carlb_brax_set_system: table has %d floats, expected %d", n_floats, TABLE_FLOATS);
    return CARLB_ERR_INVALID;
  }
  int L, nq, nqd, A;
  static_facts(env->kind, L, nq, nqd, A);
  if ((int)table[H_N_LINKS] != L || (int)table[H_N_Q] != nq || (int)table[H_N_QD] != nqd || (int)table[H_N_ACT] != A ||
      (int)table[H_N_POINTS] > MAX_POINTS) {
    set_error("carlb_brax_set_system: table does not describe env kind %d (links %d, q %d, qd %d, act %d)", env->kind,
              (int)table[H_N_LINKS], (int)table[H_N_Q], (int)table[H_N_QD], (int)table[H_N_ACT]);
    return CARLB_ERR_INVALID;
  }
  // the kernels are picked by the table's env id: it has to be the one of the handle's kind
  const int env_id = (int)table[H_ENV];
  const int want_env = env->kind - KIND_BRAX_ANT;  // CARLB_BRAX_* and EnvId enumerate the bodies in the same order
  if (env_id != want_env) {
    set_error("carlb_brax_set_system: table is for env id %d, the handle's kind %d needs %d", env_id, env->kind, want_env);
    return CARLB_ERR_INVALID;
  }
  const bool humanoid = env_id == ENV_HUMANOID || env_id == ENV_HUMANOIDSTANDUP;
  for (int l = 0; l < L; ++l) {
    const int parent = (int)table[OFF_LINKS + LINK_STRIDE * l + L_PARENT];
    const int type = (int)table[OFF_LINKS + LINK_STRIDE * l + L_TYPE];
    if (parent >= l) {
      set_error("carlb_brax_set_system: link %d has parent %d (parents must precede children)", l, parent);
      return CARLB_ERR_INVALID;
    }
    // joint types a kernel flavour does not compile in would silently be treated as plain hinges
    const bool slide = type == TYPE_SLIDE || type == TYPE_SLIDE2, stacked = type == TYPE_HINGE2 || type == TYPE_HINGE3;
    if (type < TYPE_FREE || type > TYPE_HINGE3 || (slide && !is_special_env(env_id)) || (stacked && !humanoid)) {
      set_error("carlb_brax_set_system: link %d has joint type %d, which the kernels of env id %d do not build", l, type, env_id);
      return CARLB_ERR_INVALID;
    }
  }
  const int P = (int)table[H_N_POINTS], n_pairs = (int)table[OFF_PAIR + X_N_PAIRS];
  for (int p = 0; p < P; ++p) {
    const int pl = (int)table[OFF_POINTS + POINT_STRIDE * p], slot = (int)table[OFF_POINTS + POINT_STRIDE * p + P_SCHED];
    if (pl < 0 || pl >= L || slot < 0 || slot >= P) {
      set_error("carlb_brax_set_system: contact candidate %d names link %d / schedule slot %d", p, pl, slot);
      return CARLB_ERR_INVALID;
    }
  }
  if (n_pairs < 0 || n_pairs > MAX_PAIRS || (n_pairs > 0 && env_id != ENV_PUSHER)) {
    set_error("carlb_brax_set_system: %d body-vs-body pairs (at most %d, pusher only)", n_pairs, MAX_PAIRS);
    return CARLB_ERR_INVALID;
  }
  for (int k = 0; k < n_pairs; ++k) {
    const float* pr = table + OFF_PAIR + PAIR_HEADER + PAIR_STRIDE * k;
    const int la = (int)pr[R_LINK_A], lb = (int)pr[R_LINK_B], ra = (int)pr[R_ROW_A], rb = (int)pr[R_ROW_B];
    if (la < 0 || la >= L || lb < 0 || lb >= L || ra < 0 || ra >= P || rb < 0 || rb >= P || ra == rb) {
      set_error("carlb_brax_set_system: pair %d names links %d / %d and candidate rows %d / %d", k, la, lb, ra, rb);
      return CARLB_ERR_INVALID;
    }
  }
  memcpy(h->host_table, table, sizeof(h->host_table));
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  CARLB_CUDA_CHECK(cudaMemcpy(h->dev_sys->table, table, sizeof(h->host_table), cudaMemcpyHostToDevice));
  h->have_table = true;
  h->stock_contact = stock_contact;
  return CARLB_OK;
}

static int make_brax_seg(const carlb_env* env, BraxSeg& s, const char* what, int gather_launch = -1) {
  const BraxHandle* h = static_cast<const BraxHandle*>(env->brax_sys);
  if (h == nullptr || !h->have_table) {
    set_error("%s: carlb_brax_set_system() has not been called", what);
    return CARLB_ERR_STATE;
  }
  carlb_env_info_t info;
  brax_query(env->kind, &info);
  s = BraxSeg{};
  s.n = env->n;
  s.max_steps = env->max_steps;
  s.autoreset = env->autoreset;
  s.state_words = info.state_words;
  s.obs_dim = info.obs_dim;
  s.n_ctx = info.n_param_rows;
  s.act_dim = info.act_dim;
  s.global_offset = env->global_offset;
  s.seed = h->seed;
  s.reset_rng = h->reset_rng;
  s.batch = h->batch;
  s.sys = h->dev_sys->table;
  s.state = static_cast<float*>(env->bufs.state);
  s.ctx = static_cast<const float*>(env->bufs.ctx);
  s.elapsed = env->bufs.elapsed;
  s.episode = env->bufs.rng;
  s.obs = env->bufs.obs;
  s.reward = env->bufs.reward;
  s.terminated = env->bufs.terminated;
  s.truncated = env->bufs.truncated;
  s.final_obs = env->bufs.final_obs;
  s.first_state = static_cast<float*>(env->bufs.first_state);
  s.first_obs = env->bufs.first_obs;
  s.gth = GatherDev{};
  if (env->gather != nullptr && gather_launch >= 0) gather_fill(env->gather, &s.gth, gather_launch);
  return CARLB_OK;
}

static inline int brax_grid(int n, int envs_per_cta) { return (n + envs_per_cta - 1) / envs_per_cta; }

// Envs packed per warp for the step / rollout kernels: the largest E in {4, 3, 1} whose 32/E lanes
// hold the body's links and actuators (Halfcheetah 7 links / 6 actuators and Hopper 4 / 3 -> E = 4;
// Ant 9 / 8 -> E = 3). CARLB_BRAX_PACK=1 forces one env per warp, 4 (or 3) forces packing whatever the batch size;
// unset / 0 is the automatic choice below.
static int brax_pack(const float* table, int n) {
  static const int forced = [] {
    const char* e = getenv("CARLB_BRAX_PACK");
    return e != nullptr ? atoi(e) : 0;
  }();
  int need = max((int)table[H_N_LINKS], (int)table[H_N_ACT]);
  if ((int)table[H_ENV] == ENV_HUMANOID || (int)table[H_ENV] == ENV_HUMANOIDSTANDUP) need = (int)table[H_N_LINKS];  // actions are looped
  if (forced == 1) return 1;
  // Small batches of the larger bodies are latency bound (one warp walks its substeps in ~40 us whatever the
  // grid): one env per warp spreads the contact candidates over all 32 lanes (one pass instead of three) and
  // quadruples the resident warps -- measured 10-25 % faster at <= 1 024 envs (profiles/r01m_brax_pack_sweep_*),
  // bit-identical to the packed mapping (test_packed_lanes_equal_one_env_per_warp).
  if (forced == 0 && n <= 1024 && (int)table[H_N_LINKS] >= 7) return 1;
  if (need <= Lanes<4>::LPE && forced != 3) return 4;
  if (need <= Lanes<3>::LPE) return 3;
  if (need <= Lanes<2>::LPE) return 2;
  return 1;
}

template <int W, int E, int M>
static cudaError_t launch_brax_step_we(const BraxSeg& seg, int n, const float* actions, int n_steps, uint64_t policy_seed,
                                       uint32_t step_base, const carlb_traj_t& tj, int stock_contact, cudaStream_t st) {
  typedef SmemLayoutT<W, E, M> Smem;
  static bool configured = false;
  if (!configured) {  // > 48 KB of dynamic shared memory needs the opt-in attribute
    cudaError_t e = cudaFuncSetAttribute(brax_step_kernel<W, E, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  brax_step_kernel<W, E, M><<<brax_grid(n, W * E), W * 32, sizeof(Smem), st>>>(seg, actions, n_steps, policy_seed, step_base, tj,
                                                                          stock_contact);
  return cudaGetLastError();
}

static cudaError_t launch_brax_step(const carlb_env* env, const BraxSeg& seg, const float* actions, int n_steps,
                                    uint64_t policy_seed, uint32_t step_base, const carlb_traj_t& tj, cudaStream_t st) {
  const BraxHandle* h = static_cast<const BraxHandle*>(env->brax_sys);
  const int n = env->n, sc = h->stock_contact;
  const int env_kind = (int)h->host_table[H_ENV];
  if (env_kind == ENV_HUMANOID || env_kind == ENV_HUMANOIDSTANDUP) {
    // 11 links: two envs per warp (16 lanes each; the 17 actions and 29 contact candidates go in two passes)
    if (brax_pack(h->host_table, n) == 1) return launch_brax_step_we<4, 1, MODE_HUMANOID>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
    return launch_brax_step_we<4, 2, MODE_HUMANOID>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
  }
  if (env_kind == ENV_PUSHER) {  // 9 links, slide joints, body-vs-body pairs: three envs per warp
    if (brax_pack(h->host_table, n) == 1) return launch_brax_step_we<4, 1, MODE_PUSHER>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
    return launch_brax_step_we<4, 3, MODE_PUSHER>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
  }
  if (env_kind >= ENV_INVERTED_PENDULUM) {
    // inverted pendulums / reacher (2-3 links): the instantiation with slide joints and their env layers
    if (brax_pack(h->host_table, n) == 1) return launch_brax_step_we<4, 1, MODE_SPECIAL>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
    return launch_brax_step_we<4, 4, MODE_SPECIAL>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
  }
  switch (brax_pack(h->host_table, n)) {
    case 4: return launch_brax_step_we<4, 4, MODE_LOCO>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
    case 3: return launch_brax_step_we<4, 3, MODE_LOCO>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
    default: break;
  }
  // one env per warp: 7-warp CTAs (2 per SM) keep large grids close to whole waves of 148 SMs
  if (n >= 4096) return launch_brax_step_we<7, 1, MODE_LOCO>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
  return launch_brax_step_we<4, 1, MODE_LOCO>(seg, n, actions, n_steps, policy_seed, step_base, tj, sc, st);
}

// ------------------------------------------------------------------------ goal epilogue
// BraxWalkerGoalWrapper.step (brax_walker_goal_wrapper.py:124-140), one thread per env, float64 like the
// reference's NumPy scalars: dead-reckoned xy position from two observation entries, progress reward,
// radius termination. np.linalg.norm of a 2-vector = sqrt(x*x + y*y) (this unit is built with -fmad=false).
__global__ void __launch_bounds__(128) brax_goal_kernel(int n, int obs_dim, const float* obs, int idx0, int idx1, double dt,
                                                        double* position, const double* goal, const double* radius,
                                                        double* reward, uint8_t* terminated, uint8_t* success) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 pos = reinterpret_cast<const double2*>(position)[i];
  const double2 g = reinterpret_cast<const double2*>(goal)[i];
  const float* o = obs + (size_t)i * obs_dim;
  // `np.array([state[i0], state[i1]]) * self.dt` is a float32 array times a scalar: the product is rounded
  // to float32 BEFORE it is added to the float64 position (NumPy keeps the array's dtype)
  const float dtf = (float)dt;
  const double nx = pos.x + (double)(o[idx0] * dtf), ny = pos.y + (double)(o[idx1] * dtf);
  const double cx = g.x - nx, cy = g.y - ny, px = g.x - pos.x, py = g.y - pos.y;
  const double cur = sqrt(cx * cx + cy * cy), prev = sqrt(px * px + py * py);
  const double d = prev - cur;
  reward[i] = d > 0.0 ? d : 0.0;
  reinterpret_cast<double2*>(position)[i] = make_double2(nx, ny);
  const bool reached = fabs(cur) <= radius[i];
  success[i] = reached ? 1 : 0;
  if (reached) terminated[i] = 1;
}

int brax_goal_step(const carlb_env* env, int idx0, int idx1, double dt, double* position, const double* goal,
                   const double* radius, double* reward, uint8_t* success, cudaStream_t st) {
  carlb_env_info_t info;
  brax_query(env->kind, &info);
  brax_goal_kernel<<<(env->n + 127) / 128, 128, 0, st>>>(env->n, info.obs_dim, env->bufs.obs, idx0, idx1, dt, position, goal,
                                                         radius, reward, env->bufs.terminated, success);
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

int brax_seed(const carlb_env* env, uint64_t seed, cudaStream_t st) {
  BraxHandle* h = static_cast<BraxHandle*>(env->brax_sys);
  h->seed = seed;
  CARLB_CUDA_CHECK(cudaMemsetAsync(env->bufs.rng, 0, sizeof(uint64_t) * 4 * (size_t)env->n, st));
  return CARLB_OK;
}

int brax_set_reset_rng(carlb_env* env, int mode, long long n_global) {
  BraxHandle* h = static_cast<BraxHandle*>(env->brax_sys);
  h->reset_rng = mode;
  h->batch = (uint32_t)(n_global > 0 ? n_global : 1);
  return CARLB_OK;
}

int brax_reset_from(const carlb_env* env, const uint8_t* mask, const float* q, const float* qd, cudaStream_t st) {
  BraxSeg seg;
  int rc = make_brax_seg(env, seg, "carlb_env_reset", GL_RESET);
  if (rc != CARLB_OK) return rc;
  const int env_kind = (int)static_cast<const BraxHandle*>(env->brax_sys)->host_table[H_ENV];
  if (env_kind == ENV_HUMANOID || env_kind == ENV_HUMANOIDSTANDUP)
    brax_reset_kernel<MODE_HUMANOID><<<brax_grid(env->n, 4), 128, sizeof(SmemLayoutT<4, 1, MODE_HUMANOID>), st>>>(seg, mask, q, qd);
  else if (env_kind == ENV_PUSHER)
    brax_reset_kernel<MODE_PUSHER><<<brax_grid(env->n, 4), 128, sizeof(SmemLayoutT<4, 1, MODE_PUSHER>), st>>>(seg, mask, q, qd);
  else
    brax_reset_kernel<MODE_SPECIAL><<<brax_grid(env->n, 4), 128, sizeof(SmemLayoutT<4, 1, MODE_SPECIAL>), st>>>(seg, mask, q, qd);
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

int brax_reset(const carlb_env* env, const uint8_t* mask, cudaStream_t st) {
  return brax_reset_from(env, mask, nullptr, nullptr, st);
}

int brax_step(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm) {
  (void)act_dtype;
  BraxSeg seg;
  int rc = make_brax_seg(env, seg, "carlb_env_step", GL_BRAX);
  if (rc != CARLB_OK) return rc;
  if (hm != nullptr) {
    seg.host_obs = hm->obs; seg.host_reward = hm->reward; seg.host_terminated = hm->terminated;
    seg.host_truncated = hm->truncated;
  }
  carlb_traj_t tj{};
  CARLB_CUDA_CHECK(launch_brax_step(env, seg, static_cast<const float*>(actions), 1, 0, 0, tj, st));
  g_launches++;
  return CARLB_OK;
}

int brax_rollout(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                 int act_dtype, const carlb_traj_t* traj, cudaStream_t st) {
  (void)act_dtype;
  BraxSeg seg;
  if (n_steps == 0) return CARLB_OK;
  int rc = make_brax_seg(env, seg, "carlb_env_rollout", GL_BRAX);
  if (rc != CARLB_OK) return rc;
  carlb_traj_t tj{};
  if (traj != nullptr) tj = *traj;
  seg.final_obs = nullptr;
  CARLB_CUDA_CHECK(launch_brax_step(env, seg, static_cast<const float*>(actions), n_steps, policy_seed, step_base, tj, st));
  g_launches++;
  return CARLB_OK;
}

}  // namespace CARLB_BRAX_VARIANT

// the entry points engine.h declares
#ifndef CARLB_BRAX_FMA_BUILD
int brax_query(int kind, carlb_env_info_t* out) { return strict_variant::brax_query(kind, out); }
int brax_create(carlb_env* env) { return strict_variant::brax_create(env); }
void brax_destroy(carlb_env* env) { strict_variant::brax_destroy(env); }
int brax_seed(const carlb_env* env, uint64_t seed, cudaStream_t st) { return strict_variant::brax_seed(env, seed, st); }
int brax_reset(const carlb_env* env, const uint8_t* mask, cudaStream_t st) { return strict_variant::brax_reset(env, mask, st); }
int brax_step(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm) {
  return strict_variant::brax_step(env, actions, act_dtype, st, hm);
}
int brax_rollout(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                 int act_dtype, const carlb_traj_t* traj, cudaStream_t st) {
  return strict_variant::brax_rollout(env, n_steps, policy_seed, step_base, actions, act_dtype, traj, st);
}
int brax_set_system(carlb_env* env, const float* table, int n_floats, int stock_contact) {
  return strict_variant::brax_set_system(env, table, n_floats, stock_contact);
}
int brax_reset_from(const carlb_env* env, const uint8_t* mask, const float* q, const float* qd, cudaStream_t st) {
  return strict_variant::brax_reset_from(env, mask, q, qd, st);
}
int brax_goal_step(const carlb_env* env, int idx0, int idx1, double dt, double* position, const double* goal,
                   const double* radius, double* reward, uint8_t* success, cudaStream_t st) {
  return strict_variant::brax_goal_step(env, idx0, idx1, dt, position, goal, radius, reward, success, st);
}
int brax_set_reset_rng(carlb_env* env, int mode, long long n_global) { return strict_variant::brax_set_reset_rng(env, mode, n_global); }
#else
int brax_step_fma(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm) {
  return fma_variant::brax_step(env, actions, act_dtype, st, hm);
}
int brax_rollout_fma(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                     int act_dtype, const carlb_traj_t* traj, cudaStream_t st) {
  return fma_variant::brax_rollout(env, n_steps, policy_seed, step_base, actions, act_dtype, traj, st);
}
#endif
}  // namespace carlb
