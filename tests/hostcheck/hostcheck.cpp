// TEST-ONLY shim: compiles the product's __host__ __device__ physics / RNG headers with g++ so
// the no-GPU test suite can check the kernel *source logic* against the oracle and numpy before
// GPU time is spent. Never loaded by the product (carl_b200 fails loudly without its CUDA lib).
#include <stdint.h>
#include <string.h>

#include "../../carl_b200/csrc/physics_classic.h"

using namespace carlb;

extern "C" {

void hc_pcg64_seed(uint64_t entropy, uint64_t out[4]) {
  Pcg64 g;
  pcg64_seed_from_int(g, entropy);
  out[0] = g.state_hi; out[1] = g.state_lo; out[2] = g.inc_hi; out[3] = g.inc_lo;
}

void hc_pcg64_doubles(uint64_t st[4], int n, double* out) {
  Pcg64 g{st[0], st[1], st[2], st[3]};
  for (int i = 0; i < n; ++i) out[i] = pcg64_next_double(g);
  st[0] = g.state_hi; st[1] = g.state_lo; st[2] = g.inc_hi; st[3] = g.inc_lo;
}

void hc_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
  Philox4 r = philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
  memcpy(out, r.v, sizeof(r.v));
}

// threshold predicate exactness: float fast path vs the double comparison it replaces
int hc_above_below_exact(const float* xs, int n, double thr, int* n_checked) {
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    const float x = xs[i];
    if (above(x, thr) != ((double)x > thr)) ++bad;
    if (below(x, -thr) != ((double)x < -thr)) ++bad;
    if (above(x, -thr) != ((double)x > -thr)) ++bad;
    if (below(x, thr) != ((double)x < thr)) ++bad;
  }
  *n_checked = 4 * n;
  return bad;
}

void hc_policy_action(int kind, uint64_t seed, uint64_t env_id, uint32_t step, int* ai, float* af) {
  Action a;
  PolicyStream ps = policy_stream(seed, env_id);
  switch (kind) {
    case KIND_CARTPOLE: a = policy_action<KIND_CARTPOLE>(ps, step); break;
    case KIND_PENDULUM: a = policy_action<KIND_PENDULUM>(ps, step); break;
    case KIND_ACROBOT: a = policy_action<KIND_ACROBOT>(ps, step); break;
    case KIND_MOUNTAINCAR: a = policy_action<KIND_MOUNTAINCAR>(ps, step); break;
    default: a = policy_action<KIND_MOUNTAINCAR_CONT>(ps, step); break;
  }
  *ai = a.i; *af = a.f;
}

}  // extern "C"

// state: T[n][S] ; params: T[P][n] ; rng: u64[4][n]
template <int KIND, typename T>
static void step_all(int n, T* state, const T* params, const int* ai, const float* af, uint64_t* rng, uint8_t* sbt,
                     int* elapsed, int max_steps, int autoreset, float* obs, float* reward, uint8_t* term,
                     uint8_t* trunc, float* final_obs) {
  typedef Traits<KIND> Tr;
  for (int i = 0; i < n; ++i) {
    T p[Tr::P];
    for (int r = 0; r < Tr::P; ++r) p[r] = params[(size_t)r * n + i];
    T* s = state + (size_t)i * Tr::S;
    Pcg64 g{rng[0 * n + i], rng[1 * n + i], rng[2 * n + i], rng[3 * n + i]};
    Action a{ai ? ai[i] : 0, af ? af[i] : 0.0f};
    T noise = 0;
    if (KIND == KIND_ACROBOT && p[AC_NOISE] > (T)0)
      noise = (T)pcg64_uniform(g, -(double)p[AC_NOISE], (double)p[AC_NOISE]);
    float o[8];
    StepOut so = env_step<KIND, T>(s, p, a, noise, sbt[i], o);
    elapsed[i] += 1;
    bool tr = max_steps > 0 && elapsed[i] >= max_steps;
    if (autoreset && (so.terminated || tr)) {
      if (final_obs) for (int k = 0; k < Tr::D; ++k) final_obs[(size_t)i * Tr::D + k] = o[k];
      pcg64_skip<Tr::GYM_DRAWS>(g);
      env_reset<KIND, T>(s, p, g, o);
      elapsed[i] = 0;
      sbt[i] = 0;
    }
    for (int k = 0; k < Tr::D; ++k) obs[(size_t)i * Tr::D + k] = o[k];
    reward[i] = so.reward; term[i] = so.terminated; trunc[i] = tr;
    rng[0 * n + i] = g.state_hi; rng[1 * n + i] = g.state_lo; rng[2 * n + i] = g.inc_hi; rng[3 * n + i] = g.inc_lo;
  }
}

template <int KIND, typename T>
static void reset_all(int n, T* state, const T* params, uint64_t* rng, float* obs) {
  typedef Traits<KIND> Tr;
  for (int i = 0; i < n; ++i) {
    T p[Tr::P];
    for (int r = 0; r < Tr::P; ++r) p[r] = params[(size_t)r * n + i];
    Pcg64 g{rng[0 * n + i], rng[1 * n + i], rng[2 * n + i], rng[3 * n + i]};
    pcg64_skip<Tr::GYM_DRAWS>(g);
    float o[8];
    env_reset<KIND, T>(state + (size_t)i * Tr::S, p, g, o);
    for (int k = 0; k < Tr::D; ++k) obs[(size_t)i * Tr::D + k] = o[k];
    rng[0 * n + i] = g.state_hi; rng[1 * n + i] = g.state_lo; rng[2 * n + i] = g.inc_hi; rng[3 * n + i] = g.inc_lo;
  }
}


extern "C" {

int hc_step(int kind, int f64, int n, void* state, const void* params, const int* ai, const float* af, uint64_t* rng,
            uint8_t* sbt, int* elapsed, int max_steps, int autoreset, float* obs, float* reward, uint8_t* term,
            uint8_t* trunc, float* final_obs) {
#define ARGS_F (n, (float*)state, (const float*)params, ai, af, rng, sbt, elapsed, max_steps, autoreset, obs, reward, term, trunc, final_obs)
#define ARGS_D (n, (double*)state, (const double*)params, ai, af, rng, sbt, elapsed, max_steps, autoreset, obs, reward, term, trunc, final_obs)
  switch (kind * 2 + f64) {
    case KIND_CARTPOLE * 2: step_all<KIND_CARTPOLE, float> ARGS_F; break;
    case KIND_CARTPOLE * 2 + 1: step_all<KIND_CARTPOLE, double> ARGS_D; break;
    case KIND_PENDULUM * 2: step_all<KIND_PENDULUM, float> ARGS_F; break;
    case KIND_PENDULUM * 2 + 1: step_all<KIND_PENDULUM, double> ARGS_D; break;
    case KIND_ACROBOT * 2: step_all<KIND_ACROBOT, float> ARGS_F; break;
    case KIND_ACROBOT * 2 + 1: step_all<KIND_ACROBOT, double> ARGS_D; break;
    case KIND_MOUNTAINCAR * 2: step_all<KIND_MOUNTAINCAR, float> ARGS_F; break;
    case KIND_MOUNTAINCAR * 2 + 1: step_all<KIND_MOUNTAINCAR, double> ARGS_D; break;
    case KIND_MOUNTAINCAR_CONT * 2: step_all<KIND_MOUNTAINCAR_CONT, float> ARGS_F; break;
    case KIND_MOUNTAINCAR_CONT * 2 + 1: step_all<KIND_MOUNTAINCAR_CONT, double> ARGS_D; break;
    default: return -1;
  }
  return 0;
}

int hc_reset(int kind, int f64, int n, void* state, const void* params, uint64_t* rng, float* obs) {
  switch (kind * 2 + f64) {
    case KIND_CARTPOLE * 2: reset_all<KIND_CARTPOLE, float>(n, (float*)state, (const float*)params, rng, obs); break;
    case KIND_CARTPOLE * 2 + 1: reset_all<KIND_CARTPOLE, double>(n, (double*)state, (const double*)params, rng, obs); break;
    case KIND_PENDULUM * 2: reset_all<KIND_PENDULUM, float>(n, (float*)state, (const float*)params, rng, obs); break;
    case KIND_PENDULUM * 2 + 1: reset_all<KIND_PENDULUM, double>(n, (double*)state, (const double*)params, rng, obs); break;
    case KIND_ACROBOT * 2: reset_all<KIND_ACROBOT, float>(n, (float*)state, (const float*)params, rng, obs); break;
    case KIND_ACROBOT * 2 + 1: reset_all<KIND_ACROBOT, double>(n, (double*)state, (const double*)params, rng, obs); break;
    case KIND_MOUNTAINCAR * 2: reset_all<KIND_MOUNTAINCAR, float>(n, (float*)state, (const float*)params, rng, obs); break;
    case KIND_MOUNTAINCAR * 2 + 1: reset_all<KIND_MOUNTAINCAR, double>(n, (double*)state, (const double*)params, rng, obs); break;
    case KIND_MOUNTAINCAR_CONT * 2: reset_all<KIND_MOUNTAINCAR_CONT, float>(n, (float*)state, (const float*)params, rng, obs); break;
    case KIND_MOUNTAINCAR_CONT * 2 + 1: reset_all<KIND_MOUNTAINCAR_CONT, double>(n, (double*)state, (const double*)params, rng, obs); break;
    default: return -1;
  }
  return 0;
}

// Binary-action policy through ONE stream object over an arbitrary list of steps (the fused rollout
// walks consecutive steps; the shift-register fast path must agree with a fresh stream per step).
void hc_policy_sequence(uint64_t seed, uint64_t env_id, const uint32_t* steps, int n, int* out) {
  PolicyStream ps = policy_stream(seed, env_id);
  for (int k = 0; k < n; ++k) out[k] = policy_action<KIND_CARTPOLE>(ps, steps[k]).i;
}

}  // extern "C"

// ---- JAX threefry PRNG (rng.h) compiled by g++: checked against oracle/jax_prng.py and the published known answers
extern "C" {
void hc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t out[2]) {
  carlb::threefry2x32(carlb::JaxKey{k0, k1}, x0, x1, &out[0], &out[1]);
}
void hc_jax_split(uint32_t k0, uint32_t k1, uint32_t num, uint32_t* out /* [num][2] */) {
  for (uint32_t i = 0; i < num; ++i) {
    const carlb::JaxKey k = carlb::jax_split(carlb::JaxKey{k0, k1}, num, i);
    out[2 * i] = k.k0; out[2 * i + 1] = k.k1;
  }
}
void hc_jax_uniform(uint32_t k0, uint32_t k1, uint32_t n, float lo, float hi, float* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = carlb::jax_uniform(carlb::JaxKey{k0, k1}, n, i, lo, hi);
}
void hc_jax_normal(uint32_t k0, uint32_t k1, uint32_t n, float* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = carlb::jax_normal(carlb::JaxKey{k0, k1}, n, i);
}
void hc_jax_env_reset_key(uint64_t seed, uint32_t n_resets, uint32_t batch, uint32_t env_index, uint32_t out[2]) {
  const carlb::JaxKey k = carlb::jax_env_reset_key(seed, n_resets, batch, env_index);
  out[0] = k.k0; out[1] = k.k1;
}
}  // extern "C"
