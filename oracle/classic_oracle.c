/*
 * TEST INFRASTRUCTURE -- CPU oracle for the classic-control hot path. Not part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * What it restates (float64, one env instance at a time, exactly the operation order of the
 * Python the reference executes):
 *   carl/envs/carl_env.py:321-342            CARLEnv.step -> inner gymnasium step
 *   carl/envs/gymnasium/carl_gymnasium_env.py:75-77   context injected by bare setattr
 *   carl/envs/gymnasium/classic_control/carl_{cartpole,pendulum,acrobot,mountaincar,
 *        mountaincarcontinuous}.py           feature tables (the ctx column order used below)
 * The step bodies themselves live in the third-party dependency gymnasium<1.0.0
 * (pyproject.toml:35; last matching release 0.29.1), which is NOT vendored in the reference tree:
 * gymnasium/envs/classic_control/{cartpole,pendulum,acrobot,mountain_car,
 * continuous_mountain_car}.py and gymnasium/wrappers/time_limit.py are restated from the
 * published sources.
 *
 * Pinning: CartPole is pinned by the published gymnasium known answers (seed-0 reset vector,
 * action-1 step) in tests/golden/gymnasium_known_answers.json. Pendulum / Acrobot / MountainCar
 * have no golden vectors anywhere in the reference => PARITY UNPINNED for those three.
 *
 * ctx layout: double ctx[n][F], F and column order = the env's get_context_features() order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

/* gymnasium TimeLimit.step: elapsed += 1; truncated = elapsed >= max_episode_steps */
static unsigned char time_limit(int *elapsed, int max_steps) {
  *elapsed += 1;
  return (unsigned char)(max_steps > 0 && *elapsed >= max_steps);
}

/* ------------------------------------------------------------------ CartPole (A.1)
 * ctx columns: gravity, masscart, masspole, length, force_mag, tau, initial_state_lower,
 * initial_state_upper (carl_cartpole.py:16-42).
 * applied_mode == 0 reproduces the reference: CartPoleEnv caches total_mass = 1.1 and
 * polemass_length = 0.05 in __init__ and setattr never refreshes them. */
void oracle_cartpole_step(int n, double *state, const double *ctx, const int *action, int *elapsed,
                          unsigned char *sbt, int max_steps, int applied_mode, float *obs, double *reward,
                          unsigned char *term, unsigned char *trunc) {
  const double theta_threshold_radians = 12 * 2 * PI / 360;
  const double x_threshold = 2.4;
  for (int i = 0; i < n; ++i) {
    double *s = state + 4 * i;
    const double *c = ctx + 8 * i;
    const double gravity = c[0], masscart = c[1], masspole = c[2], length = c[3], force_mag = c[4], tau = c[5];
    double total_mass = 0.1 + 1.0, polemass_length = 0.1 * 0.5;
    if (applied_mode) {
      total_mass = masspole + masscart;
      polemass_length = masspole * length;
    }
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    double force = action[i] == 1 ? force_mag : -force_mag;
    double costheta = cos(theta), sintheta = sin(theta);
    double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
    double thetaacc =
        (gravity * sintheta - costheta * temp) / (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
    double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
    x = x + tau * x_dot;
    x_dot = x_dot + tau * xacc;
    theta = theta + tau * theta_dot;
    theta_dot = theta_dot + tau * thetaacc;
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
    int terminated =
        x < -x_threshold || x > x_threshold || theta < -theta_threshold_radians || theta > theta_threshold_radians;
    double r;
    if (!terminated) {
      r = 1.0;
    } else if (sbt[i] == 0) { /* steps_beyond_terminated is None */
      sbt[i] = 1;
      r = 1.0;
    } else {
      r = 0.0;
    }
    reward[i] = r;
    term[i] = (unsigned char)terminated;
    trunc[i] = time_limit(&elapsed[i], max_steps);
    for (int k = 0; k < 4; ++k) obs[4 * i + k] = (float)s[k];
  }
}

/* ------------------------------------------------------------------ Pendulum (A.2)
 * ctx columns: gravity(dead), dt, g, m, l, initial_angle_max, initial_velocity_max
 * (carl_pendulum.py:16-39). action is float32 (clipped to +-2, then used as float64). */
static double py_mod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0) {
    if ((b < 0) != (r < 0)) r += b;
  } else {
    r = copysign(0.0, b);
  }
  return r;
}

void oracle_pendulum_step(int n, double *state, const double *ctx, const float *action, int *elapsed, int max_steps,
                          float *obs, double *reward, unsigned char *term, unsigned char *trunc) {
  const double max_speed = 8, max_torque = 2.0;
  for (int i = 0; i < n; ++i) {
    double *s = state + 2 * i;
    const double *c = ctx + 7 * i;
    const double dt = c[1], g = c[2], m = c[3], l = c[4];
    double th = s[0], thdot = s[1];
    float uf = action[i];
    if (uf < (float)-max_torque) uf = (float)-max_torque;
    if (uf > (float)max_torque) uf = (float)max_torque;
    double u = (double)uf;
    double an = py_mod(th + PI, 2 * PI) - PI;
    double costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u);
    double newthdot = thdot + (3 * g / (2 * l) * sin(th) + 3.0 / (m * (l * l)) * u) * dt;
    if (newthdot < -max_speed) newthdot = -max_speed;
    if (newthdot > max_speed) newthdot = max_speed;
    double newth = th + newthdot * dt;
    s[0] = newth;
    s[1] = newthdot;
    obs[3 * i + 0] = (float)cos(newth);
    obs[3 * i + 1] = (float)sin(newth);
    obs[3 * i + 2] = (float)newthdot;
    reward[i] = -costs;
    term[i] = 0;
    trunc[i] = time_limit(&elapsed[i], max_steps);
  }
}

/* ------------------------------------------------------------------- Acrobot (A.3)
 * ctx columns: LINK_LENGTH_1, LINK_LENGTH_2, LINK_MASS_1, LINK_MASS_2, LINK_COM_POS_1,
 * LINK_COM_POS_2, LINK_MOI, MAX_VEL_1, MAX_VEL_2, torque_noise_max, INITIAL_ANGLE_LOWER,
 * INITIAL_ANGLE_UPPER, INITIAL_VELOCITY_LOWER, INITIAL_VELOCITY_UPPER (carl_acrobot.py:16-69).
 * noise[i] is the torque noise the env RNG produced for this step (0 when disabled). */
static void acrobot_dsdt(const double *sa, const double *c, double *out) {
  const double m1 = c[2], m2 = c[3], l1 = c[0], lc1 = c[4], lc2 = c[5], I1 = c[6], I2 = c[6];
  const double g = 9.8;
  const double a = sa[4];
  const double theta1 = sa[0], theta2 = sa[1], dtheta1 = sa[2], dtheta2 = sa[3];
  double d1 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2;
  double d2 = m2 * (lc2 * lc2 + l1 * lc2 * cos(theta2)) + I2;
  double phi2 = m2 * lc2 * g * cos(theta1 + theta2 - PI / 2.0);
  double phi1 = -m2 * l1 * lc2 * (dtheta2 * dtheta2) * sin(theta2) - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2) +
                (m1 * lc1 + m2 * l1) * g * cos(theta1 - PI / 2) + phi2;
  double ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * (dtheta1 * dtheta1) * sin(theta2) - phi2) /
                    (m2 * (lc2 * lc2) + I2 - (d2 * d2) / d1);
  double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
  out[0] = dtheta1; out[1] = dtheta2; out[2] = ddtheta1; out[3] = ddtheta2; out[4] = 0.0;
}

static double wrap(double x, double m, double M) {
  double diff = M - m;
  while (x > M) x = x - diff;
  while (x < m) x = x + diff;
  return x;
}
static double bound(double x, double m, double M) { return fmin(fmax(x, m), M); }

void oracle_acrobot_step(int n, double *state, const double *ctx, const int *action, const double *noise, int *elapsed,
                         int max_steps, float *obs, double *reward, unsigned char *term, unsigned char *trunc) {
  static const double AVAIL_TORQUE[3] = {-1.0, 0.0, +1};
  const double dt = 0.2;
  for (int i = 0; i < n; ++i) {
    double *s = state + 4 * i;
    const double *c = ctx + 14 * i;
    double torque = AVAIL_TORQUE[action[i]];
    if (c[9] > 0) torque += noise ? noise[i] : 0.0;
    double y0[5] = {s[0], s[1], s[2], s[3], torque};
    double k1[5], k2[5], k3[5], k4[5], y[5];
    const double dt2 = dt / 2.0;
    acrobot_dsdt(y0, c, k1);
    for (int k = 0; k < 5; ++k) y[k] = y0[k] + dt2 * k1[k];
    acrobot_dsdt(y, c, k2);
    for (int k = 0; k < 5; ++k) y[k] = y0[k] + dt2 * k2[k];
    acrobot_dsdt(y, c, k3);
    for (int k = 0; k < 5; ++k) y[k] = y0[k] + dt * k3[k];
    acrobot_dsdt(y, c, k4);
    double ns[4];
    for (int k = 0; k < 4; ++k) ns[k] = y0[k] + dt / 6.0 * (k1[k] + 2 * k2[k] + 2 * k3[k] + k4[k]);
    ns[0] = wrap(ns[0], -PI, PI);
    ns[1] = wrap(ns[1], -PI, PI);
    ns[2] = bound(ns[2], -c[7], c[7]);
    ns[3] = bound(ns[3], -c[8], c[8]);
    for (int k = 0; k < 4; ++k) s[k] = ns[k];
    int terminated = (-cos(ns[0]) - cos(ns[1] + ns[0])) > 1.0;
    reward[i] = terminated ? 0.0 : -1.0;
    term[i] = (unsigned char)terminated;
    trunc[i] = time_limit(&elapsed[i], max_steps);
    obs[6 * i + 0] = (float)cos(ns[0]); obs[6 * i + 1] = (float)sin(ns[0]);
    obs[6 * i + 2] = (float)cos(ns[1]); obs[6 * i + 3] = (float)sin(ns[1]);
    obs[6 * i + 4] = (float)ns[2]; obs[6 * i + 5] = (float)ns[3];
  }
}

/* --------------------------------------------------------------- MountainCar (A.4)
 * ctx columns: min_position, max_position, max_speed, goal_position, goal_velocity, force,
 * gravity, min_position_start, max_position_start, min_velocity_start, max_velocity_start
 * (carl_mountaincar.py:16-51). */
void oracle_mountaincar_step(int n, double *state, const double *ctx, const int *action, int *elapsed, int max_steps,
                             float *obs, double *reward, unsigned char *term, unsigned char *trunc) {
  for (int i = 0; i < n; ++i) {
    double *s = state + 2 * i;
    const double *c = ctx + 11 * i;
    double position = s[0], velocity = s[1];
    velocity += (action[i] - 1) * c[5] + cos(3 * position) * (-c[6]);
    velocity = fmin(fmax(velocity, -c[2]), c[2]); /* np.clip */
    position += velocity;
    position = fmin(fmax(position, c[0]), c[1]);
    if (position == c[0] && velocity < 0) velocity = 0;
    int terminated = position >= c[3] && velocity >= c[4];
    s[0] = position;
    s[1] = velocity;
    reward[i] = -1.0;
    term[i] = (unsigned char)terminated;
    trunc[i] = time_limit(&elapsed[i], max_steps);
    obs[2 * i] = (float)position;
    obs[2 * i + 1] = (float)velocity;
  }
}

/* ----------------------------------------------------- MountainCarContinuous (A.5)
 * ctx columns: min_position, max_position, max_speed, goal_position, goal_velocity, power,
 * min_position_start, max_position_start, min_velocity_start, max_velocity_start
 * (carl_mountaincarcontinuous.py:16-48). State is stored as float32 after every step. */
void oracle_mountaincar_cont_step(int n, double *state, const double *ctx, const float *action, int *elapsed,
                                  int max_steps, float *obs, double *reward, unsigned char *term, unsigned char *trunc) {
  for (int i = 0; i < n; ++i) {
    double *s = state + 2 * i;
    const double *c = ctx + 10 * i;
    double position = s[0], velocity = s[1];
    double a0 = (double)action[i];
    double force = fmin(fmax(a0, -1.0), 1.0);
    velocity += force * c[5] - 0.0025 * cos(3 * position);
    if (velocity > c[2]) velocity = c[2];
    if (velocity < -c[2]) velocity = -c[2];
    position += velocity;
    if (position > c[1]) position = c[1];
    if (position < c[0]) position = c[0];
    if (position == c[0] && velocity < 0) velocity = 0;
    int terminated = position >= c[3] && velocity >= c[4];
    double r = 0;
    if (terminated) r = 100.0;
    r -= pow(a0, 2) * 0.1;
    s[0] = (double)(float)position;
    s[1] = (double)(float)velocity;
    reward[i] = r;
    term[i] = (unsigned char)terminated;
    trunc[i] = time_limit(&elapsed[i], max_steps);
    obs[2 * i] = (float)position;
    obs[2 * i + 1] = (float)velocity;
  }
}

/* ------------------------------------------------------------------ CPU baseline leg
 * Random-policy CartPole stepping over n env instances for `steps` steps with episode restarts
 * (state re-drawn U(lower, upper) from a per-env xorshift stream -- timing only, not parity),
 * parallelised over envs with OpenMP when available. Returns the number of env-steps done. */
static inline uint64_t xorshift64(uint64_t *s) {
  uint64_t x = *s;
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  return *s = x;
}

long long oracle_cartpole_rollout_baseline(int n, int steps, double *state, const double *ctx, int max_steps,
                                           int applied_mode, uint64_t seed, int n_threads, double *return_sum) {
  double total = 0.0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(static) reduction(+ : total)
#endif
  for (int i = 0; i < n; ++i) {
    uint64_t rs = seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i * 0xD1B54A32D192ED03ULL + 1ULL;
    int elapsed = 0;
    unsigned char sbt = 0, term, trunc;
    float obs[4];
    double r;
    for (int t = 0; t < steps; ++t) {
      int a = (int)(xorshift64(&rs) >> 63);
      oracle_cartpole_step(1, state + 4 * i, ctx + 8 * i, &a, &elapsed, &sbt, max_steps, applied_mode, obs, &r, &term,
                           &trunc);
      total += r;
      if (term || trunc) {
        const double lo = ctx[8 * i + 6], hi = ctx[8 * i + 7];
        for (int k = 0; k < 4; ++k)
          state[4 * i + k] = lo + (hi - lo) * ((double)(xorshift64(&rs) >> 11) * (1.0 / 9007199254740992.0));
        elapsed = 0;
        sbt = 0;
      }
    }
  }
  if (return_sum) *return_sum = total;
  return (long long)n * (long long)steps;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
