/*
 * TEST INFRASTRUCTURE -- CPU oracle for the Brax-locomotion hot path (spring backend). Not part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load it.
 *
 * PARITY UNPINNED. The arithmetic of this path lives in the third-party packages brax==0.12.1
 * and jax (pyproject.toml:62-63), which are absent from the reference tree and not installable
 * here, and the reference's own tests pin no numeric result for it (test/test_brax_env.py:8-23 is
 * construct + reset). This file restates, from the published Brax v2 sources,
 * what the reference reaches through
 *   carl/envs/carl_env.py:321-342  CARLEnv.step
 *   carl/envs/brax/wrappers.py:62-67,74-78   jitted step(state, action) -> obs, reward, done
 *   brax.envs.{ant,half_cheetah,hopper}.step -> PipelineEnv.pipeline_step (n_frames x
 *   brax.spring.pipeline.step) + EpisodeWrapper + AutoResetWrapper,
 * and pipeline_init / kinematics.inverse for reset and the observation, with the context changes
 * of carl/envs/brax/carl_brax_env.py:255-292 (gravity, ang_damping, link masses, friction,
 * elasticity) arriving as per-env scalars. Every assumption is listed in DESIGN.md §brax; the
 * system constants come from the packed table of carl_b200/envs/brax_system.py (a data format,
 * shared by design). float32 throughout like the reference's JAX pipeline; serial loops, one env
 * at a time.
 *
 * Conventions: quaternion (w,x,y,z). Per-link state row (13 floats): COM pos[3], link rot[4],
 * COM vel[3], angular vel[3].
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- table layout (must match carl_b200/envs/brax_system.py) ---- */
enum { MAXL = 12, MAXP = 32, MAXQ = 24, HDR = 32, LSTR = 40, PSTR = 8, DSTR = 16, MAXOBS = 256, MAXPAIR = 4, RSTR = 16, RHDR = 8 };
enum { OFF_L = HDR, OFF_P = HDR + LSTR * MAXL, OFF_Q = HDR + LSTR * MAXL + PSTR * MAXP, OFF_D = HDR + LSTR * MAXL + PSTR * MAXP + MAXQ,
       OFF_R = HDR + LSTR * MAXL + PSTR * MAXP + MAXQ + DSTR * MAXL,
       TABLE_N = HDR + LSTR * MAXL + PSTR * MAXP + MAXQ + DSTR * MAXL + RHDR + RSTR * MAXPAIR };
/* pair region (body-vs-body contacts, pusher): header, then rows: capsule side (link, candidate row, segment ends,
 * radius), sphere side (link, candidate row, centre, radius) */
enum { xN_PAIRS = 0, xPLANE_Z, xOBS_LINK0, xOBS_LINK1, xOBS_LINK2 };
enum { rLINK_A = 0, rROW_A = 1, rA0 = 2, rA1 = 5, rRAD_A = 8, rLINK_B = 9, rROW_B = 10, rB0 = 11, rRAD_B = 14 };
/* dof rows (stacked hinges): actuator / gear / range of dofs 1 and 2, sign of each coordinate against the joint frame */
enum { dACT1 = 0, dACT2, dGEAR1, dGEAR2, dLO1, dHI1, dLO2, dHI2, dSIGN0, dSIGN1, dSIGN2 };
enum { hN_LINKS = 0, hN_Q, hN_QD, hN_POINTS, hN_FRAMES, hDT, hENV, hN_ACT, hK, hCV, hKL, hCA, hERP, hVDAMP, hMSCALE,
       hISCALE, hNOISE, hCTRL, hHEALTHY, hZMIN, hZMAX, hFWD, hAMIN, hAMAX, hEXCL, hQDCLIP, hTERM, hMAXCP, hQDUNI,
       hSITE_LINK, hQDNOISE, hACTSCALE };
enum { lPARENT = 0, lTYPE, lQ, lQD, lTPOS = 4, lTROT = 7, lJPOS = 11, lJROT = 14, lLO = 18, lHI = 19, lCOM = 20,
       lIROT = 23, lIDIAG = 27, lMASS = 30, lGEAR = 31, lACT = 32, lCLO = 33, lCHI = 34, lFIRSTP = 35, lNP = 36, lSITE = 37 };
/* T_SLIDE: prismatic joint along the joint x axis (carts of brax.envs.inverted_pendulum /
 * inverted_double_pendulum); T_SLIDE2: two prismatic dofs, joint x and y (target of brax.envs.reacher) */
/* T_HINGE2 / T_HINGE3: two / three stacked revolute dofs (universal / spherical joints of brax humanoid.xml):
 * the joint rotation is Rx(a0) Ry(a1) Rz(a2) in the joint frame whose x, y are the first two MJCF axes */
enum { T_FREE = 0, T_HINGE = 1, T_SLIDE = 2, T_PLANAR = 3, T_SLIDE2 = 4, T_HINGE2 = 5, T_HINGE3 = 6 };
enum { E_ANT = 0, E_CHEETAH = 1, E_HOPPER = 2, E_WALKER2D = 3, E_IPENDULUM = 4, E_IDPENDULUM = 5, E_REACHER = 6,
       E_HUMANOID = 7, E_STANDUP = 8, E_PUSHER = 9 };

/* Arithmetic type of the restatement: float (the reference's JAX pipeline) by default; built a
 * second time with -DORACLE_F64 as the round-off-free yardstick that tells float32 noise (stiff
 * springs amplify it) from algorithmic differences. Tables / contexts / actions stay float32. */
#ifdef ORACLE_F64
typedef double real;
#define NAME(x) x##64
#define SQRT sqrt
#define POW pow
#define EXP exp
#define ATAN2 atan2
#define SIN sin
#define COS cos
#define FMIN fmin
#define FMAX fmax
#else
typedef float real;
#define NAME(x) x
#define SQRT sqrtf
#define POW powf
#define EXP expf
#define ATAN2 atan2f
#define SIN sinf
#define COS cosf
#define FMIN fminf
#define FMAX fmaxf
#endif
typedef real f3[3];
typedef real f4[4];

static void cross3(const real *a, const real *b, real *o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static real dot3(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void qmul(const real *a, const real *b, real *o) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
static void qconj(const real *a, real *o) { o[0] = a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = -a[3]; }
static void qnorm(real *q) {
  real n = SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  real inv = 1.0f / n;
  for (int k = 0; k < 4; ++k) q[k] *= inv;
}
/* brax.math.rotate */
static void rot3(const real *v, const real *q, real *o) {
  const real *u = q + 1;
  real s = q[0];
  real uv = dot3(u, v), uu = dot3(u, u);
  f3 c;
  cross3(u, v, c);
  for (int k = 0; k < 3; ++k) o[k] = 2.0f * uv * u[k] + (s * s - uu) * v[k] + 2.0f * s * c[k];
}
static void irot3(const real *v, const real *q, real *o) {
  f4 qc;
  qconj(q, qc);
  rot3(v, qc, o);
}

/* link-frame origin and its velocity from the COM row */
static void origin_of(const real *row, const real *lt, real *o) {
  f3 rc;
  rot3(lt + lCOM, row + 3, rc);
  for (int k = 0; k < 3; ++k) o[k] = row[k] - rc[k];
}
static void origin_vel_of(const real *row, const real *lt, real *o) {
  f3 rc, w;
  rot3(lt + lCOM, row + 3, rc);
  cross3(row + 10, rc, w);
  for (int k = 0; k < 3; ++k) o[k] = row[7 + k] - w[k];
}
static real eff_mass(real m, const real *sys) { return POW(m, 1.0f - sys[hMSCALE]); }
static void apply_inv_inertia(const real *v, const real *rot, const real *lt, const real *sys, real *o) {
  f4 r;
  f3 w;
  qmul(rot, lt + lIROT, r);
  irot3(v, r, w);
  real e = 1.0f - sys[hISCALE];
  for (int k = 0; k < 3; ++k) w[k] *= 1.0f / POW(lt[lIDIAG + k], e);
  rot3(w, r, o);
}

/* joint frame quantities of link l against its parent (kinematics.world_to_joint) */
typedef struct {
  f3 ac_pos, ap_pos, jpos, jvel, jang;
  f4 ap_rot, jrot;
  real psi;
} jframe_t;

/* Stacked hinges (kinematics.axis_angle_ang restated as intrinsic x-y'-z'' Euler angles): the joint rotation
 * R = Rx(a0) Ry(a1) Rz(a2); torque / rate axes in the joint frame: e_x, the line of nodes Rx(a0) e_y, and the
 * child's z axis R e_z. */
static void euler_axes(const real *jrot, real *ang, real ax[3][3]) {
  f3 ex = {1, 0, 0}, ey = {0, 1, 0}, ez = {0, 0, 1}, xc, yc, zc;
  rot3(ex, jrot, xc);
  rot3(ey, jrot, yc);
  rot3(ez, jrot, zc);
  real c1 = SQRT(zc[1] * zc[1] + zc[2] * zc[2]);
  real inv = 1.0f / (1e-10f + c1);
  ang[0] = ATAN2(-zc[1], zc[2]);
  ang[1] = ATAN2(zc[0], c1);
  ang[2] = ATAN2(-yc[0], xc[0]);
  ax[0][0] = 1; ax[0][1] = 0; ax[0][2] = 0;
  ax[1][0] = 0; ax[1][1] = zc[2] * inv; ax[1][2] = -zc[1] * inv;
  for (int k = 0; k < 3; ++k) ax[2][k] = zc[k];
}
static int type_dofs(int type) { return type == T_HINGE3 || type == T_PLANAR ? 3 : (type == T_HINGE2 || type == T_SLIDE2 ? 2 : 1); }
/* range and sign of dof k of a stacked hinge */
static void dof_range(const real *sys, int l, int k, real *lo, real *hi, real *sign) {
  const real *lt = sys + OFF_L + LSTR * l, *dt = sys + OFF_D + DSTR * l;
  *sign = dt[dSIGN0 + k];
  if (k == 0) { *lo = lt[lLO]; *hi = lt[lHI]; }
  else if (k == 1) { *lo = dt[dLO1]; *hi = dt[dHI1]; }
  else { *lo = dt[dLO2]; *hi = dt[dHI2]; }
}

static void joint_frame(const real *sys, const real *rows, int l, jframe_t *j) {
  const real *lt = sys + OFF_L + LSTR * l;
  const real *c = rows + 13 * l;
  int parent = (int)lt[lPARENT];
  f3 xc, xp = {0, 0, 0}, vp = {0, 0, 0}, wp = {0, 0, 0}, tmp, tmp2;
  f4 xprot = {1, 0, 0, 0}, ac_rot, t;
  origin_of(c, lt, xc);
  rot3(lt + lJPOS, c + 3, tmp);
  for (int k = 0; k < 3; ++k) j->ac_pos[k] = xc[k] + tmp[k];
  qmul(c + 3, lt + lJROT, ac_rot);
  const real *p = 0;
  if (parent >= 0) {
    p = rows + 13 * parent;
    origin_of(p, sys + OFF_L + LSTR * parent, xp);
    memcpy(xprot, p + 3, sizeof(f4));
    memcpy(wp, p + 10, sizeof(f3));
  }
  rot3(lt + lJPOS, lt + lTROT, tmp);
  for (int k = 0; k < 3; ++k) tmp[k] += lt[lTPOS + k];
  rot3(tmp, xprot, tmp2);
  for (int k = 0; k < 3; ++k) j->ap_pos[k] = xp[k] + tmp2[k];
  qmul(xprot, lt + lTROT, t);
  qmul(t, lt + lJROT, j->ap_rot);
  if (parent >= 0) {
    for (int k = 0; k < 3; ++k) tmp[k] = j->ap_pos[k] - p[k];
    cross3(p + 10, tmp, tmp2);
    for (int k = 0; k < 3; ++k) vp[k] = p[7 + k] + tmp2[k];
  }
  f3 vc;
  for (int k = 0; k < 3; ++k) tmp[k] = j->ac_pos[k] - c[k];
  cross3(c + 10, tmp, tmp2);
  for (int k = 0; k < 3; ++k) vc[k] = c[7 + k] + tmp2[k];
  for (int k = 0; k < 3; ++k) tmp[k] = j->ac_pos[k] - j->ap_pos[k];
  irot3(tmp, j->ap_rot, j->jpos);
  f4 apc;
  qconj(j->ap_rot, apc);
  qmul(apc, ac_rot, j->jrot);
  for (int k = 0; k < 3; ++k) tmp[k] = vc[k] - vp[k];
  irot3(tmp, j->ap_rot, j->jvel);
  for (int k = 0; k < 3; ++k) tmp[k] = c[10 + k] - wp[k];
  irot3(tmp, j->ap_rot, j->jang);
  f3 ey = {0, 1, 0}, yc;
  rot3(ey, j->jrot, yc);
  j->psi = ATAN2(yc[2], yc[1]);
}

/* one spring substep over all links of one env (brax.spring.pipeline.step) */
static void substep(const real *sys, real *rows, const real *ctx, real (*tau)[3]) {
  const int L = (int)sys[hN_LINKS], P = (int)sys[hN_POINTS];
  const real dt = sys[hDT];
  const real gravity = ctx[0], friction = ctx[1], elasticity = ctx[2], ang_damping = ctx[3];
  real F[MAXL][3], T[MAXL][3];
  memset(F, 0, sizeof(F));
  memset(T, 0, sizeof(T));
  /* 1. joint spring / damper / limit / actuator forces (spring/joints.py resolve) */
  real pF[MAXL][3], pT[MAXL][3];
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l;
    int type = (int)lt[lTYPE], parent = (int)lt[lPARENT];
    memset(pF[l], 0, sizeof(f3));
    memset(pT[l], 0, sizeof(f3));
    if (type == T_FREE) continue;
    jframe_t j;
    joint_frame(sys, rows, l, &j);
    const real k = sys[hK] * ctx[4], cv = sys[hCV], kl = sys[hKL], ca = sys[hCA];
    f3 ex = {1, 0, 0}, fv, fa, axc;
    rot3(ex, j.jrot, axc);
    cross3(axc, ex, fa);
    for (int q = 0; q < 3; ++q) fa[q] *= k;
    if (type == T_PLANAR) {
      fv[0] = -k * j.jpos[0] - cv * j.jvel[0]; fv[1] = 0; fv[2] = 0;
      fa[1] -= ca * j.jang[1];
      fa[2] -= ca * j.jang[2];
    } else if (type == T_HINGE2 || type == T_HINGE3) {
      /* universal / spherical: position spring on the anchor, no axis-alignment torque; the universal joint keeps
       * the child's y axis perpendicular to the parent's x axis (third Euler angle = 0) with a spring torque;
       * range limit and actuator torque per dof about its Euler axis; damping on the whole relative rate */
      real ang[3], ax[3][3];
      euler_axes(j.jrot, ang, ax);
      for (int q = 0; q < 3; ++q) fv[q] = -k * j.jpos[q] - cv * j.jvel[q];
      fa[0] = fa[1] = fa[2] = 0.0f;
      if (type == T_HINGE2) {
        f3 ey = {0, 1, 0}, yc, proj, t2;
        rot3(ey, j.jrot, yc);
        real inv = 1.0f / (1e-10f + SQRT(yc[1] * yc[1] + yc[2] * yc[2]));
        proj[0] = 0; proj[1] = yc[1] * inv; proj[2] = yc[2] * inv;
        cross3(yc, proj, t2);
        for (int q = 0; q < 3; ++q) fa[q] += k * t2[q];
      }
      int nd = type_dofs(type);
      for (int d = 0; d < nd; ++d) {
        real lo, hi, sg;
        dof_range(sys, l, d, &lo, &hi, &sg);
        real coord = sg * ang[d], dang = 0.0f;
        if (coord < lo) dang = lo - coord;
        if (coord > hi) dang = hi - coord;
        real tq = sg * (kl * dang + tau[l][d]);
        for (int q = 0; q < 3; ++q) fa[q] += tq * ax[d][q];
      }
      for (int q = 0; q < 3; ++q) fa[q] -= ca * j.jang[q];
    } else if (type == T_SLIDE || type == T_SLIDE2) {
      /* prismatic: no relative rotation at all (a second alignment torque on the y axes), springs on the
       * constrained offsets only, range limit and actuator force along the first sliding axis */
      f3 ey = {0, 1, 0}, ayc, t2;
      rot3(ey, j.jrot, ayc);
      cross3(ayc, ey, t2);
      for (int q = 0; q < 3; ++q) fa[q] += k * t2[q];
      for (int q = 0; q < 3; ++q) fa[q] -= ca * j.jang[q];
      real dpos = 0.0f;
      if (j.jpos[0] < lt[lLO]) dpos = lt[lLO] - j.jpos[0];
      if (j.jpos[0] > lt[lHI]) dpos = lt[lHI] - j.jpos[0];
      fv[0] = kl * dpos + tau[l][0];
      fv[1] = (type == T_SLIDE) ? (-k * j.jpos[1] - cv * j.jvel[1]) : 0.0f;
      fv[2] = -k * j.jpos[2] - cv * j.jvel[2];
    } else {
      for (int q = 0; q < 3; ++q) fv[q] = -k * j.jpos[q] - cv * j.jvel[q];
      real dang = 0.0f;
      if (j.psi < lt[lLO]) dang = lt[lLO] - j.psi;
      if (j.psi > lt[lHI]) dang = lt[lHI] - j.psi;
      fa[0] += kl * dang;
      for (int q = 0; q < 3; ++q) fa[q] -= ca * j.jang[q];
      fa[0] += tau[l][0];
    }
    f3 Fw, Tw, r, rxF;
    rot3(fv, j.ap_rot, Fw);
    rot3(fa, j.ap_rot, Tw);
    const real *c = rows + 13 * l;
    for (int q = 0; q < 3; ++q) r[q] = j.ac_pos[q] - c[q];
    cross3(r, Fw, rxF);
    for (int q = 0; q < 3; ++q) { F[l][q] += Fw[q]; T[l][q] += Tw[q] + rxF[q]; }
    if (parent >= 0) {
      const real *p = rows + 13 * parent;
      for (int q = 0; q < 3; ++q) r[q] = j.ap_pos[q] - p[q];
      cross3(r, Fw, rxF);
      for (int q = 0; q < 3; ++q) { pF[l][q] = -Fw[q]; pT[l][q] = -Tw[q] - rxF[q]; }
    }
  }
  for (int l = 0; l < L; ++l)        /* reactions onto the parents, children in index order */
    for (int c = l + 1; c < L; ++c)
      if ((int)sys[OFF_L + LSTR * c + lPARENT] == l)
        for (int q = 0; q < 3; ++q) { F[l][q] += pF[c][q]; T[l][q] += pT[c][q]; }
  /* 2. semi-implicit velocity update (integrator.integrate_xdd) */
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l;
    real *s = rows + 13 * l;
    real m = eff_mass(ctx[5 + l], sys);
    f3 alpha;
    apply_inv_inertia(T[l], s + 3, lt, sys, alpha);
    real inv_m = 1.0f / m;
    f3 acc = {F[l][0] * inv_m, F[l][1] * inv_m, gravity + F[l][2] * inv_m};
    real dv = EXP(sys[hVDAMP] * dt), da = EXP(ang_damping * dt);
    for (int q = 0; q < 3; ++q) {
      s[7 + q] = (s[7 + q] + acc[q] * dt) * dv;
      s[10 + q] = (s[10 + q] + alpha[q] * dt) * da;
    }
  }
  /* 3. ground contacts (spring/collisions.py), impulses averaged per link over active contacts */
  real ps[MAXL][3], ts[MAXL][3], na[MAXL];
  memset(ps, 0, sizeof(ps));
  memset(ts, 0, sizeof(ts));
  memset(na, 0, sizeof(na));
  for (int p = 0; p < P; ++p) {
    const real *pt = sys + OFF_P + PSTR * p;
    int l = (int)pt[0];
    const real *lt = sys + OFF_L + LSTR * l;
    const real *s = rows + 13 * l;
    real fr = friction < 0.0f ? pt[5] : friction;
    real el = elasticity < 0.0f ? pt[6] : elasticity;
    f3 org, c, loc;
    origin_of(s, lt, org);
    rot3(pt + 1, s + 3, loc);
    for (int q = 0; q < 3; ++q) c[q] = org[q] + loc[q];
    if (pt[4] < 0.0f) continue; /* a row that only receives the impulse of a body-vs-body pair */
    real dist = c[2] - pt[4], pen = -dist;
    if (!(pen > 0.0f)) continue;
    f3 n = {0, 0, 1}, cpos = {c[0], c[1], 0.5f * dist}, rel, rv, tmp;
    for (int q = 0; q < 3; ++q) rel[q] = cpos[q] - s[q];
    cross3(s + 10, rel, tmp);
    for (int q = 0; q < 3; ++q) rv[q] = s[7 + q] + tmp[q];
    real nv = dot3(n, rv);
    real inv_m = 1.0f / eff_mass(ctx[5 + l], sys);
    f3 rxn, t1, t2;
    cross3(rel, n, rxn);
    apply_inv_inertia(rxn, s + 3, lt, sys, t1);
    cross3(t1, rel, t2);
    real ang = dot3(n, t2);
    real bvel = sys[hERP] * pen / sys[hDT];
    real imp = (-1.0f * (1.0f + el) * nv + bvel) / (inv_m + ang);
    f3 vd;
    for (int q = 0; q < 3; ++q) vd[q] = rv[q] - nv * n[q];
    real sd = SQRT(dot3(vd, vd));
    real impd = sd / (inv_m + ang);
    real inv_sd = 1.0f / (1e-6f + sd);
    impd = FMIN(impd, fr * imp);
    int apply_n = (nv < 0.0f) && (imp > 0.0f);
    int apply_d = apply_n && (sd > 0.01f);
    if (!apply_n) continue;
    f3 tot = {0, 0, imp};
    if (apply_d)
      for (int q = 0; q < 3; ++q) tot[q] += -impd * (vd[q] * inv_sd);
    cross3(rel, tot, tmp);
    for (int q = 0; q < 3; ++q) { ps[l][q] += tot[q]; ts[l][q] += tmp[q]; }
    na[l] += 1.0f;
  }
  /* 3b. body-vs-body contacts (spring/collisions.py with two dynamic bodies): capsule A against sphere B. The contact
   * normal points from B's centre to the closest point of A's segment (it pushes A away from B), the contact point
   * lies midway between the two surfaces; the impulse divides by both inverse masses and both angular terms and acts
   * with opposite signs on the two links; each side counts as one active contact of its link */
  const int NPAIR = (int)sys[OFF_R + xN_PAIRS];
  for (int k = 0; k < NPAIR; ++k) {
    const real *pr = sys + OFF_R + RHDR + RSTR * k;
    int la = (int)pr[rLINK_A], lb = (int)pr[rLINK_B];
    const real *lta = sys + OFF_L + LSTR * la, *ltb = sys + OFF_L + LSTR * lb;
    const real *sa = rows + 13 * la, *sb = rows + 13 * lb;
    const real *pta = sys + OFF_P + PSTR * (int)pr[rROW_A];
    real fr = friction < 0.0f ? pta[5] : friction;
    real el = elasticity < 0.0f ? pta[6] : elasticity;
    f3 oa, ob, a0, a1, cb, t0;
    origin_of(sa, lta, oa);
    origin_of(sb, ltb, ob);
    rot3(pr + rA0, sa + 3, t0);
    for (int q = 0; q < 3; ++q) a0[q] = oa[q] + t0[q];
    rot3(pr + rA1, sa + 3, t0);
    for (int q = 0; q < 3; ++q) a1[q] = oa[q] + t0[q];
    rot3(pr + rB0, sb + 3, t0);
    for (int q = 0; q < 3; ++q) cb[q] = ob[q] + t0[q];
    /* closest point of the segment a0-a1 to the sphere centre */
    f3 ab, ac, cp, dvec;
    for (int q = 0; q < 3; ++q) { ab[q] = a1[q] - a0[q]; ac[q] = cb[q] - a0[q]; }
    real tt = dot3(ac, ab) / dot3(ab, ab);
    tt = FMIN(FMAX(tt, 0.0f), 1.0f);
    for (int q = 0; q < 3; ++q) { cp[q] = a0[q] + tt * ab[q]; dvec[q] = cp[q] - cb[q]; }
    real dist = SQRT(dot3(dvec, dvec));
    real pen = pr[rRAD_A] + pr[rRAD_B] - dist;
    if (!(pen > 0.0f)) continue;
    f3 n, cpos, ra, rb, va, vb, cv_, tmp;
    real inv_d = 1.0f / (1e-6f + dist);
    for (int q = 0; q < 3; ++q) n[q] = dvec[q] * inv_d;
    /* midway between the sphere's surface point and the capsule's surface point */
    for (int q = 0; q < 3; ++q) cpos[q] = 0.5f * ((cb[q] + pr[rRAD_B] * n[q]) + (cp[q] - pr[rRAD_A] * n[q]));
    for (int q = 0; q < 3; ++q) { ra[q] = cpos[q] - sa[q]; rb[q] = cpos[q] - sb[q]; }
    cross3(sa + 10, ra, tmp);
    for (int q = 0; q < 3; ++q) va[q] = sa[7 + q] + tmp[q];
    cross3(sb + 10, rb, tmp);
    for (int q = 0; q < 3; ++q) vb[q] = sb[7 + q] + tmp[q];
    for (int q = 0; q < 3; ++q) cv_[q] = va[q] - vb[q];
    real nv = dot3(n, cv_);
    real inv_ma = 1.0f / eff_mass(ctx[5 + la], sys), inv_mb = 1.0f / eff_mass(ctx[5 + lb], sys);
    f3 rxn, t1, t2a, t2b;
    cross3(ra, n, rxn);
    apply_inv_inertia(rxn, sa + 3, lta, sys, t1);
    cross3(t1, ra, t2a);
    cross3(rb, n, rxn);
    apply_inv_inertia(rxn, sb + 3, ltb, sys, t1);
    cross3(t1, rb, t2b);
    real ang = dot3(n, t2a) + dot3(n, t2b);
    real denom = inv_ma + inv_mb + ang;
    real bvel = sys[hERP] * pen / sys[hDT];
    real imp = (-1.0f * (1.0f + el) * nv + bvel) / denom;
    f3 vd;
    for (int q = 0; q < 3; ++q) vd[q] = cv_[q] - nv * n[q];
    real sd = SQRT(dot3(vd, vd));
    real impd = sd / denom;
    real inv_sd = 1.0f / (1e-6f + sd);
    impd = FMIN(impd, fr * imp);
    int apply_n = (nv < 0.0f) && (imp > 0.0f);
    int apply_d = apply_n && (sd > 0.01f);
    if (!apply_n) continue;
    f3 tot;
    for (int q = 0; q < 3; ++q) tot[q] = imp * n[q];
    if (apply_d)
      for (int q = 0; q < 3; ++q) tot[q] += -impd * (vd[q] * inv_sd);
    cross3(ra, tot, tmp);
    for (int q = 0; q < 3; ++q) { ps[la][q] += tot[q]; ts[la][q] += tmp[q]; }
    na[la] += 1.0f;
    cross3(rb, tot, tmp);
    for (int q = 0; q < 3; ++q) { ps[lb][q] -= tot[q]; ts[lb][q] -= tmp[q]; }
    na[lb] += 1.0f;
  }
  /* 4. delta-velocity + pose integration */
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l;
    real *s = rows + 13 * l;
    if (na[l] > 0.0f) {
      real inv_n = 1.0f / na[l];
      real sc = inv_n / eff_mass(ctx[5 + l], sys);
      f3 tt = {ts[l][0] * inv_n, ts[l][1] * inv_n, ts[l][2] * inv_n}, dw;
      apply_inv_inertia(tt, s + 3, lt, sys, dw);
      for (int q = 0; q < 3; ++q) { s[7 + q] += ps[l][q] * sc; s[10 + q] += dw[q]; }
    }
    for (int q = 0; q < 3; ++q) s[q] += s[7 + q] * dt;
    f4 w = {0.0f, s[10] * 0.5f * dt, s[11] * 0.5f * dt, s[12] * 0.5f * dt}, d;
    qmul(w, s + 3, d);
    for (int q = 0; q < 4; ++q) s[3 + q] += d[q];
    qnorm(s + 3);
  }
}

/* kinematics.inverse: generalized coordinates of one env */
static void inverse_kinematics(const real *sys, const real *rows, real *q, real *qd) {
  const int L = (int)sys[hN_LINKS];
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l;
    const real *s = rows + 13 * l;
    int type = (int)lt[lTYPE], qi = (int)lt[lQ], qdi = (int)lt[lQD];
    if (type == T_FREE) {
      f3 o, vo, al;
      origin_of(s, lt, o);
      origin_vel_of(s, lt, vo);
      irot3(s + 10, s + 3, al);
      for (int k = 0; k < 3; ++k) { q[qi + k] = o[k]; qd[qdi + k] = vo[k]; qd[qdi + 3 + k] = al[k]; }
      for (int k = 0; k < 4; ++k) q[qi + 3 + k] = s[3 + k];
    } else {
      jframe_t j;
      joint_frame(sys, rows, l, &j);
      if (type == T_PLANAR) {
        f3 o, vo;
        origin_of(s, lt, o);
        origin_vel_of(s, lt, vo);
        q[qi] = o[0] - lt[lTPOS]; q[qi + 1] = o[2] - lt[lTPOS + 2]; q[qi + 2] = j.psi;
        qd[qdi] = vo[0]; qd[qdi + 1] = vo[2]; qd[qdi + 2] = j.jang[0];
      } else if (type == T_SLIDE || type == T_SLIDE2) {
        q[qi] = j.jpos[0];
        qd[qdi] = j.jvel[0];
        if (type == T_SLIDE2) { q[qi + 1] = j.jpos[1]; qd[qdi + 1] = j.jvel[1]; }
      } else if (type == T_HINGE2 || type == T_HINGE3) {
        real ang[3], ax[3][3];
        euler_axes(j.jrot, ang, ax);
        for (int d = 0; d < type_dofs(type); ++d) {
          real sg = sys[OFF_D + DSTR * l + dSIGN0 + d];
          q[qi + d] = sg * ang[d];
          qd[qdi + d] = sg * dot3(ax[d], j.jang);
        }
      } else {
        q[qi] = j.psi;
        qd[qdi] = j.jang[0];
      }
    }
  }
}

/* world position of the env's site (pendulum tip / reacher fingertip): x.take(link).do(Transform(pos=site)) */
static void site_pos(const real *sys, const real *rows, real *o) {
  int l = (int)sys[hSITE_LINK];
  const real *lt = sys + OFF_L + LSTR * l;
  f3 org, r;
  origin_of(rows + 13 * l, lt, org);
  rot3(lt + lSITE, rows + 13 * l + 3, r);
  for (int k = 0; k < 3; ++k) o[k] = org[k] + r[k];
}

/* centre of mass of the whole body with the spring backend's effective link masses (brax humanoid._com) */
static real body_com(const real *sys, const real *rows, const real *ctx, real *com) {
  const int L = (int)sys[hN_LINKS];
  real msum = 0.0f;
  com[0] = com[1] = com[2] = 0.0f;
  for (int l = 0; l < L; ++l) {
    real m = eff_mass(ctx[5 + l], sys);
    msum += m;
    for (int k = 0; k < 3; ++k) com[k] += m * rows[13 * l + k];
  }
  for (int k = 0; k < 3; ++k) com[k] = com[k] / msum;
  return msum;
}

/* actuator torque of every actuated dof: gear * clip(action) (brax.actuator.to_tau) */
static void actuator_taus(const real *sys, const real *act, real (*tau)[3]) {
  const int L = (int)sys[hN_LINKS];
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l, *dt = sys + OFF_D + DSTR * l;
    int ai[3] = {(int)lt[lACT], (int)dt[dACT1], (int)dt[dACT2]};
    real gear[3] = {lt[lGEAR], dt[dGEAR1], dt[dGEAR2]};
    int nd = (int)lt[lTYPE] == T_FREE ? 0 : type_dofs((int)lt[lTYPE]);
    for (int d = 0; d < 3; ++d) {
      tau[l][d] = 0.0f;
      if (d < nd && ai[d] >= 0) tau[l][d] = gear[d] * FMIN(FMAX(act[ai[d]], lt[lCLO]), lt[lCHI]);
    }
  }
}

/* brax.envs.humanoid._get_obs: q[2:], qd, per link (inertia about the body COM in the world frame 3x3, mass),
 * per link (mass-weighted COM velocity, angular velocity), actuator torques in qd layout */
static void humanoid_obs(const real *sys, const real *rows, const real *q, const real *qd, const real *ctx,
                         const real *act, real *obs) {
  const int L = (int)sys[hN_LINKS], nq = (int)sys[hN_Q], nqd = (int)sys[hN_QD];
  int k = 0;
  for (int i = 2; i < nq; ++i) obs[k++] = q[i];
  for (int i = 0; i < nqd; ++i) obs[k++] = qd[i];
  f3 com;
  real msum = body_com(sys, rows, ctx, com);
  real e = 1.0f - sys[hISCALE];
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l, *s = rows + 13 * l;
    real m = eff_mass(ctx[5 + l], sys);
    f3 p = {s[0] - com[0], s[1] - com[1], s[2] - com[2]};
    f4 r;
    qmul(s + 3, lt + lIROT, r);
    real R[3][3]; /* R[a][c]: component a of the rotated basis vector c */
    for (int c = 0; c < 3; ++c) {
      f3 b = {c == 0, c == 1, c == 2}, o;
      rot3(b, r, o);
      for (int a = 0; a < 3; ++a) R[a][c] = o[a];
    }
    real ie[3];
    for (int c = 0; c < 3; ++c) ie[c] = POW(lt[lIDIAG + c], e);
    real pp = dot3(p, p);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        real rot_i = (R[a][0] * ie[0] * R[b][0] + R[a][1] * ie[1] * R[b][1]) + R[a][2] * ie[2] * R[b][2];
        real par = (a == b ? pp : 0.0f) - p[a] * p[b];
        obs[k++] = rot_i + m * par;
      }
    obs[k++] = m;
  }
  for (int l = 0; l < L; ++l) {
    const real *s = rows + 13 * l;
    real m = eff_mass(ctx[5 + l], sys);
    for (int c = 0; c < 3; ++c) obs[k++] = m * s[7 + c] / msum;
    for (int c = 0; c < 3; ++c) obs[k++] = s[10 + c];
  }
  real tau[MAXL][3];
  actuator_taus(sys, act, tau);
  for (int i = 0; i < nqd; ++i) obs[k + i] = 0.0f;
  for (int l = 0; l < L; ++l) {
    const real *lt = sys + OFF_L + LSTR * l;
    if ((int)lt[lTYPE] == T_FREE) continue;
    for (int d = 0; d < type_dofs((int)lt[lTYPE]); ++d) obs[k + (int)lt[lQD] + d] = tau[l][d];
  }
}

static void make_obs(const real *sys, const real *rows, const real *q, const real *qd, const real *ctx, const real *act,
                     real *obs) {
  int nq = (int)sys[hN_Q], nqd = (int)sys[hN_QD], ex = (int)sys[hEXCL], env = (int)sys[hENV];
  real clip = sys[hQDCLIP];
  int k = 0;
  if (env == E_HUMANOID || env == E_STANDUP) {
    humanoid_obs(sys, rows, q, qd, ctx, act, obs);
    return;
  }
  if (env == E_PUSHER) { /* brax.envs.pusher._get_obs: q[:7], qd[:7], COM of the wrist-flex link, the object, the goal */
    for (int i = 0; i < 7; ++i) obs[k++] = q[i];
    for (int i = 0; i < 7; ++i) obs[k++] = qd[i];
    for (int j = 0; j < 3; ++j) {
      const real *s = rows + 13 * (int)sys[OFF_R + xOBS_LINK0 + j];
      obs[k++] = s[0]; obs[k++] = s[1]; obs[k++] = s[2] - sys[OFF_R + xPLANE_Z];
    }
    return;
  }
  if (env == E_IDPENDULUM) { /* brax.envs.inverted_double_pendulum._get_obs */
    obs[k++] = q[0];
    obs[k++] = SIN(q[1]); obs[k++] = SIN(q[2]);
    obs[k++] = COS(q[1]); obs[k++] = COS(q[2]);
    for (int i = 0; i < 3; ++i) obs[k++] = FMIN(FMAX(qd[i], -10.0f), 10.0f);
    return;
  }
  if (env == E_REACHER) { /* brax.envs.reacher._get_obs: cos, sin, target q, arm qd, tip - target */
    f3 tip, tgt;
    site_pos(sys, rows, tip);
    origin_of(rows + 13 * 2, sys + OFF_L + LSTR * 2, tgt);
    obs[k++] = COS(q[0]); obs[k++] = COS(q[1]);
    obs[k++] = SIN(q[0]); obs[k++] = SIN(q[1]);
    obs[k++] = q[2]; obs[k++] = q[3];
    obs[k++] = qd[0]; obs[k++] = qd[1];
    for (int i = 0; i < 3; ++i) obs[k++] = tip[i] - tgt[i];
    return;
  }
  for (int i = ex; i < nq; ++i) obs[k++] = q[i];
  for (int i = 0; i < nqd; ++i) {
    real v = qd[i];
    if (clip > 0.0f) v = FMIN(FMAX(v, -clip), clip);
    obs[k++] = v;
  }
}

/* pipeline_init: forward kinematics from (q, qd) into the COM rows */
void NAME(brax_oracle_init)(const float *sys_f, int n, const float *q_all, const float *qd_all, real *state, int state_words,
                      float *obs, int obs_dim, const float *ctx, int n_ctx) {
  real sys[TABLE_N];
  for (int i = 0; i < TABLE_N; ++i) sys[i] = sys_f[i];
  const int L = (int)sys[hN_LINKS], nq = (int)sys[hN_Q], nqd = (int)sys[hN_QD];
  for (int e = 0; e < n; ++e) {
    real q[MAXQ], qd[MAXQ];
    for (int i = 0; i < nq; ++i) q[i] = q_all[(size_t)e * nq + i];
    for (int i = 0; i < nqd; ++i) qd[i] = qd_all[(size_t)e * nqd + i];
    real *rows = state + (size_t)e * state_words;
    memset(rows, 0, sizeof(real) * state_words);
    for (int l = 0; l < L; ++l) {
      const real *lt = sys + OFF_L + LSTR * l;
      int type = (int)lt[lTYPE], parent = (int)lt[lPARENT];
      const real *ql = q + (int)lt[lQ], *qdl = qd + (int)lt[lQD];
      f3 xpos, xvel, xang;
      f4 xrot;
      if (type == T_FREE) {
        for (int k = 0; k < 3; ++k) { xpos[k] = ql[k]; xvel[k] = qdl[k]; }
        for (int k = 0; k < 4; ++k) xrot[k] = ql[3 + k];
        qnorm(xrot);
        rot3(qdl + 3, xrot, xang);
      } else {
        f3 ex = {1, 0, 0}, axis, xp = {0, 0, 0}, vp = {0, 0, 0}, wp = {0, 0, 0};
        f4 xprot = {1, 0, 0, 0};
        rot3(ex, lt + lJROT, axis);
        if (parent >= 0) {
          const real *p = rows + 13 * parent;
          const real *plt = sys + OFF_L + LSTR * parent;
          origin_of(p, plt, xp);
          origin_vel_of(p, plt, vp);
          memcpy(xprot, p + 3, sizeof(f4));
          memcpy(wp, p + 10, sizeof(f3));
        }
        real angle, rate;
        f3 trans = {0, 0, 0}, tvel = {0, 0, 0};
        int stacked = (type == T_HINGE2 || type == T_HINGE3);
        if (stacked) {
          angle = 0.0f; rate = 0.0f;
        } else if (type == T_PLANAR) {
          trans[0] = ql[0]; trans[2] = ql[1];
          tvel[0] = qdl[0]; tvel[2] = qdl[1];
          angle = ql[2]; rate = qdl[2];
        } else if (type == T_SLIDE || type == T_SLIDE2) {
          for (int k = 0; k < 3; ++k) { trans[k] = axis[k] * ql[0]; tvel[k] = axis[k] * qdl[0]; }
          if (type == T_SLIDE2) {
            f3 ey = {0, 1, 0}, axis_y;
            rot3(ey, lt + lJROT, axis_y);
            for (int k = 0; k < 3; ++k) { trans[k] += axis_y[k] * ql[1]; tvel[k] += axis_y[k] * qdl[1]; }
          }
          angle = 0.0f; rate = 0.0f;
        } else {
          angle = ql[0]; rate = qdl[0];
        }
        real h = 0.5f * angle, sn = SIN(h);
        f4 jrot = {COS(h), axis[0] * sn, axis[1] * sn, axis[2] * sn};
        f3 wj = {0, 0, 0}; /* stacked hinges: relative angular velocity in the joint frame */
        if (stacked) {
          /* joint rotation Rx(a0) Ry(a1) Rz(a2) in the joint frame, carried into the link frame by the joint
           * orientation; rates about e_x, Rx(a0) e_y, Rx(a0) Ry(a1) e_z */
          const real *dof = sys + OFF_D + DSTR * l;
          f4 acc = {1, 0, 0, 0}, t, jc;
          for (int d = 0; d < type_dofs(type); ++d) {
            real a = dof[dSIGN0 + d] * ql[d], hh = 0.5f * a;
            f4 e = {COS(hh), 0, 0, 0};
            e[1 + d] = SIN(hh);
            f3 b = {d == 0, d == 1, d == 2}, bw;
            rot3(b, acc, bw);
            for (int k = 0; k < 3; ++k) wj[k] += dof[dSIGN0 + d] * qdl[d] * bw[k];
            qmul(acc, e, t);
            memcpy(acc, t, sizeof(f4));
          }
          qmul(lt + lJROT, acc, t);
          qconj(lt + lJROT, jc);
          qmul(t, jc, jrot);
        }
        qnorm(jrot);
        f3 rj, jpos, lpos, tmp, tmp2;
        rot3(lt + lJPOS, jrot, rj);
        for (int k = 0; k < 3; ++k) jpos[k] = trans[k] + (lt[lJPOS + k] - rj[k]);
        rot3(jpos, lt + lTROT, tmp);
        for (int k = 0; k < 3; ++k) lpos[k] = lt[lTPOS + k] + tmp[k];
        f4 lrot;
        qmul(lt + lTROT, jrot, lrot);
        rot3(lpos, xprot, tmp);
        for (int k = 0; k < 3; ++k) xpos[k] = xp[k] + tmp[k];
        qmul(xprot, lrot, xrot);
        qnorm(xrot);
        for (int k = 0; k < 3; ++k) tmp[k] = xpos[k] - xp[k];
        cross3(wp, tmp, tmp2);
        rot3(tvel, xprot, tmp);
        for (int k = 0; k < 3; ++k) xvel[k] = vp[k] + tmp2[k] + tmp[k];
        f3 ar = {axis[0] * rate, axis[1] * rate, axis[2] * rate};
        rot3(ar, xrot, tmp);
        if (stacked) {
          f4 xpt;
          qmul(xprot, lt + lTROT, xpt);
          rot3(wj, lt + lJROT, tmp2);
          rot3(tmp2, xpt, tmp);
        }
        for (int k = 0; k < 3; ++k) xang[k] = wp[k] + tmp[k];
        if (lt[lJPOS] != 0.0f || lt[lJPOS + 1] != 0.0f || lt[lJPOS + 2] != 0.0f) {
          /* the joint position is the pivot of the rotation: the link origin (pivot - R j_pos) swings around it at
           * -w x (R j_pos), w = the joint's own angular rate in the link-transform frame */
          f3 wl = {axis[0] * rate, axis[1] * rate, axis[2] * rate}, sw, t3;
          if (stacked) rot3(wj, lt + lJROT, wl);
          cross3(wl, rj, sw);
          for (int k = 0; k < 3; ++k) sw[k] = 0.0f - sw[k];
          rot3(sw, lt + lTROT, t3);
          rot3(t3, xprot, sw);
          for (int k = 0; k < 3; ++k) xvel[k] += sw[k];
        }
      }
      real *s = rows + 13 * l;
      f3 rc, w;
      rot3(lt + lCOM, xrot, rc);
      cross3(xang, rc, w);
      for (int k = 0; k < 3; ++k) { s[k] = xpos[k] + rc[k]; s[7 + k] = xvel[k] + w[k]; s[10 + k] = xang[k]; }
      for (int k = 0; k < 4; ++k) s[3 + k] = xrot[k];
    }
    real qq[MAXQ], qqd[MAXQ], ob[MAXOBS], c[5 + MAXL], zero_act[MAXQ];
    /* context row of the env (link masses enter the humanoid's observation); stock masses without one */
    for (int i = 0; i < 5; ++i) c[i] = 0.0f;
    for (int l = 0; l < L; ++l) c[5 + l] = sys[OFF_L + LSTR * l + lMASS];
    if (ctx != 0)
      for (int i = 0; i < n_ctx; ++i) c[i] = ctx[(size_t)e * n_ctx + i];
    memset(zero_act, 0, sizeof(zero_act));
    inverse_kinematics(sys, rows, qq, qqd);
    make_obs(sys, rows, qq, qqd, c, zero_act, ob);
    for (int i = 0; i < obs_dim; ++i) obs[(size_t)e * obs_dim + i] = (float)ob[i];
  }
}

/* One env-step of n envs: actuator torques, n_frames substeps, env layer, EpisodeWrapper,
 * AutoResetWrapper (autoreset != 0). ctx rows: gravity, friction, elasticity, ang_damping,
 * joint-stiffness scale (legacy `joint_stiffness` extension, 1 = stock), masses. */
void NAME(brax_oracle_step)(const float *sys_f, int n, real *state, int state_words, const float *ctx, int n_ctx,
                      const float *actions, int *elapsed, int max_steps, int autoreset, const real *first_state,
                      const float *first_obs, float *obs, int obs_dim, float *reward, unsigned char *done_out,
                      float *final_obs) {
  real sys[TABLE_N];
  for (int i = 0; i < TABLE_N; ++i) sys[i] = sys_f[i];
  const int A = (int)sys[hN_ACT], NF = (int)sys[hN_FRAMES], env = (int)sys[hENV];
  const int nq = (int)sys[hN_Q], nqd = (int)sys[hN_QD];
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int e = 0; e < n; ++e) {
    real *rows = state + (size_t)e * state_words;
    real c[5 + MAXL], act[MAXQ];
    for (int i = 0; i < n_ctx; ++i) c[i] = ctx[(size_t)e * n_ctx + i];
    for (int i = 0; i < A; ++i) act[i] = actions[(size_t)e * A + i];
    real tau[MAXL][3];
    real act_sq = 0.0f;
    for (int a = 0; a < A; ++a) act_sq += act[a] * act[a];
    actuator_taus(sys, act, tau);
    f3 o0, o1;
    origin_of(rows, sys + OFF_L, o0);
    if (env == E_HUMANOID) body_com(sys, rows, c, o0); /* humanoid: velocity of the body COM, not of the torso */
    real near0 = 0.0f, dist0 = 0.0f;
    if (env == E_PUSHER) { /* brax.envs.pusher.step: the reward reads the positions BEFORE the pipeline advances */
      const real *tip = rows + 13 * (int)sys[OFF_R + xOBS_LINK0], *ob_ = rows + 13 * (int)sys[OFF_R + xOBS_LINK1],
                 *gl = rows + 13 * (int)sys[OFF_R + xOBS_LINK2];
      f3 v1 = {ob_[0] - tip[0], ob_[1] - tip[1], ob_[2] - tip[2]}, v2 = {ob_[0] - gl[0], ob_[1] - gl[1], ob_[2] - gl[2]};
      near0 = SQRT(dot3(v1, v1));
      dist0 = SQRT(dot3(v2, v2));
    }
    for (int f = 0; f < NF; ++f) substep(sys, rows, c, tau);
    origin_of(rows, sys + OFF_L, o1);
    real z_root = o1[2];
    if (env == E_HUMANOID) body_com(sys, rows, c, o1);
    real q[MAXQ], qd[MAXQ];
    inverse_kinematics(sys, rows, q, qd);
    real ob[MAXOBS];
    make_obs(sys, rows, q, qd, c, act, ob);
    real dt_env = sys[hDT] * sys[hN_FRAMES];
    real xvel = (o1[0] - o0[0]) / dt_env;
    int healthy = 1;
    if (env == E_ANT) {
      healthy = !(o1[2] < sys[hZMIN]) && !(o1[2] > sys[hZMAX]);
    } else if (env == E_HUMANOID) { /* brax.envs.humanoid.step: torso z inside healthy_z_range */
      healthy = !(z_root < sys[hZMIN]) && !(z_root > sys[hZMAX]);
    } else if (env == E_HOPPER) {
      int ok = 1;
      for (int i = 2; i < nq; ++i) ok = ok && (q[i] > -100.0f) && (q[i] < 100.0f);
      for (int i = 0; i < nqd; ++i) ok = ok && (qd[i] > -100.0f) && (qd[i] < 100.0f);
      healthy = ok && (sys[hZMIN] < o1[2]) && (o1[2] < sys[hZMAX]) && (sys[hAMIN] < q[2]) && (q[2] < sys[hAMAX]);
    } else if (env == E_WALKER2D) { /* brax walker2d: z and torso angle ranges, no state-range check */
      healthy = !(o1[2] < sys[hZMIN]) && !(o1[2] > sys[hZMAX]) && !(q[2] > sys[hAMAX]) && !(q[2] < sys[hAMIN]);
    }
    real r = sys[hFWD] * xvel + sys[hHEALTHY] - sys[hCTRL] * act_sq;
    int done = (sys[hTERM] > 0.0f) && !healthy;
    if (env == E_IPENDULUM) { /* brax.envs.inverted_pendulum.step: reward 1, done = |obs[1]| > 0.2 */
      r = 1.0f;
      done = (q[1] < 0.0f ? -q[1] : q[1]) > 0.2f;
    } else if (env == E_IDPENDULUM) { /* tip = x.take(2) o (0,0,0.6); x, _, y = tip.pos */
      f3 tip;
      site_pos(sys, rows, tip);
      real dist_penalty = 0.01f * (tip[0] * tip[0]) + (tip[2] - 2.0f) * (tip[2] - 2.0f);
      real vel_penalty = 1e-3f * (qd[1] * qd[1]) + 5e-3f * (qd[2] * qd[2]);
      r = 10.0f - dist_penalty - vel_penalty;
      done = tip[2] <= 1.0f;
    } else if (env == E_PUSHER) { /* reward_dist + 0.1 reward_ctrl + 0.5 reward_near, never done */
      r = ((0.0f - dist0) + sys[hCTRL] * (0.0f - act_sq)) + 0.5f * (0.0f - near0);
      done = 0;
    } else if (env == E_STANDUP) { /* brax.envs.humanoidstandup.step: uph_cost + 1 - quad_ctrl_cost, never done */
      r = (z_root - 0.0f) / dt_env + sys[hHEALTHY] - sys[hCTRL] * act_sq;
      done = 0;
    } else if (env == E_REACHER) { /* reward_dist + reward_ctrl = -|tip - target| - sum(a^2) */
      real d2 = ob[8] * ob[8] + ob[9] * ob[9] + ob[10] * ob[10];
      r = (0.0f - SQRT(d2)) + (0.0f - act_sq);
      done = 0;
    }
    elapsed[e] += 1;
    if (max_steps > 0 && elapsed[e] >= max_steps) done = 1;
    if (done && autoreset) {
      if (final_obs)
        for (int i = 0; i < obs_dim; ++i) final_obs[(size_t)e * obs_dim + i] = (float)ob[i];
      memcpy(rows, first_state + (size_t)e * state_words, sizeof(real) * state_words);
      for (int i = 0; i < obs_dim; ++i) ob[i] = first_obs[(size_t)e * obs_dim + i];
      elapsed[e] = 0;
    }
    for (int i = 0; i < obs_dim; ++i) obs[(size_t)e * obs_dim + i] = (float)ob[i];
    reward[e] = (float)r;
    done_out[e] = (unsigned char)done;
  }
}
