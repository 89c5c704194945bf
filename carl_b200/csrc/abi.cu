// extern "C" surface of libcarlb (declared in include/carlb.h).
#include <cuda_runtime.h>
#include <limits>
#include <type_traits>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "engine.h"

namespace carlb {

std::atomic<long long> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

Segment make_segment(const carlb_env* env, int act_dtype) {
  Segment s{};
  s.kind = env->kind;
  s.n = env->n;
  s.max_steps = env->max_steps;
  s.autoreset = env->autoreset;
  s.act_dtype = act_dtype;
  s.global_offset = env->global_offset;
  s.state = env->bufs.state;
  s.ctx = env->bufs.ctx;
  s.elapsed = env->bufs.elapsed;
  s.sbt = env->bufs.sbt;
  s.rng = env->bufs.rng;
  s.obs = env->bufs.obs;
  s.reward = env->bufs.reward;
  s.terminated = env->bufs.terminated;
  s.truncated = env->bufs.truncated;
  s.final_obs = env->bufs.final_obs;
  s.gth = GatherDev{};
  s.host_obs = nullptr; s.host_reward = nullptr; s.host_terminated = nullptr; s.host_truncated = nullptr;
  return s;
}

// true when `p` is page-locked host memory the device can address directly (UVA)
static bool is_mapped_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
}

template <int KIND> static void fill_info(carlb_env_info_t* o) {
  typedef Traits<KIND> Tr;
  o->kind = KIND;
  o->state_words = Tr::S;
  o->obs_dim = Tr::D;
  o->act_dim = Tr::A;
  o->act_discrete = Tr::DISCRETE ? 1 : 0;
  o->n_actions = Tr::N_ACTIONS;
  o->n_param_rows = Tr::P;
  o->n_step_rows = Tr::P_STEP;
  o->default_max_steps = Tr::MAX_STEPS;
  o->gym_reset_draws = Tr::GYM_DRAWS;
  o->act_low = 0.0f;
  o->act_high = (float)(Tr::N_ACTIONS - 1);
  if (KIND == KIND_PENDULUM) { o->act_low = -2.0f; o->act_high = 2.0f; }
  if (KIND == KIND_MOUNTAINCAR_CONT) { o->act_low = -1.0f; o->act_high = 1.0f; }
}

static bool is_classic(int kind) { return kind >= 0 && kind < KIND_CLASSIC_COUNT; }
static bool is_brax(int kind) { return kind >= KIND_BRAX_ANT && kind <= KIND_BRAX_PUSHER; }

static bool valid_act_dtype(const carlb_env* env, int act_dtype) {
  carlb_env_info_t info;
  if (carlb_query_env(env->kind, &info) != CARLB_OK) return false;
  if (info.act_discrete) return act_dtype == CARLB_ACT_I32 || act_dtype == CARLB_ACT_I64 || act_dtype == CARLB_ACT_U8;
  return act_dtype == CARLB_ACT_F32;
}

static int check_ready(const carlb_env* env, const char* what) {
  if (env == nullptr) {
    set_error("%s: null handle", what);
    return CARLB_ERR_INVALID;
  }
  if (!env->bound) {
    set_error("%s: carlb_env_bind() has not been called", what);
    return CARLB_ERR_STATE;
  }
  return CARLB_OK;
}

}  // namespace carlb

using namespace carlb;

// Host-side staging of a caller's (pageable) action array into the page-locked block the step reads:
// ONE pass that copies and, for discrete spaces, range-checks (`assert self.action_space.contains(action)`
// of the gymnasium envs, e.g. cartpole.py step) -- replaces a min, a max and a copy on the Python side.
template <typename A>
static bool all_below(const A* __restrict__ src, int64_t count, uint64_t n_actions) {
  typedef typename std::make_unsigned<A>::type U;
  if (n_actions > (uint64_t)std::numeric_limits<U>::max()) return true;
  const U limit = (U)n_actions;
  if ((n_actions & (n_actions - 1)) == 0) {  // power of two: OR-reduce (plain SSE2), any stray bit shows
    U acc = 0;
    for (int64_t i = 0; i < count; ++i) acc |= (U)src[i];
    return acc < limit;
  }
  U bad = 0;
  for (int64_t i = 0; i < count; ++i) bad |= (U)((U)src[i] >= limit);
  return bad == 0;
}

// range check of an action array that is already page-locked (nothing to copy)
static int check_only(const void* src, int64_t count, int act_dtype, int n_actions) {
  bool ok = true;
  if (n_actions > 0) {
    switch (act_dtype) {
      case CARLB_ACT_I32: ok = all_below((const int32_t*)src, count, (uint64_t)n_actions); break;
      case CARLB_ACT_I64: ok = all_below((const int64_t*)src, count, (uint64_t)n_actions); break;
      case CARLB_ACT_U8: ok = all_below((const uint8_t*)src, count, (uint64_t)n_actions); break;
      default: break;
    }
  }
  if (!ok) {
    set_error("invalid action: values must lie in [0, %d)", n_actions);
    return CARLB_ERR_INVALID;
  }
  return CARLB_OK;
}

template <typename A>
static bool stage_checked(A* __restrict__ dst, const A* __restrict__ src, int64_t count, uint64_t n_actions) {
  typedef typename std::make_unsigned<A>::type U;  // as unsigned, a negative value exceeds every valid action
  const U limit = n_actions > (uint64_t)std::numeric_limits<U>::max() ? std::numeric_limits<U>::max() : (U)n_actions;
  const bool unbounded = n_actions > (uint64_t)std::numeric_limits<U>::max();
  if (!unbounded && (n_actions & (n_actions - 1)) == 0) {  // power of two: copy + OR-reduce
    U acc = 0;
    for (int64_t i = 0; i < count; ++i) {
      const A v = src[i];
      dst[i] = v;
      acc |= (U)v;
    }
    return acc < limit;
  }
  U bad = 0;
  for (int64_t i = 0; i < count; ++i) {
    const A v = src[i];
    dst[i] = v;
    bad |= (U)((U)v >= limit);
  }
  return unbounded || bad == 0;
}

extern "C" {

int carlb_abi_version(void) { return CARLB_ABI_VERSION; }
const char* carlb_last_error(void) { return g_err; }
int64_t carlb_launch_count(void) { return (int64_t)g_launches.load(); }

int carlb_query_env(int kind, carlb_env_info_t* out) {
  if (out == nullptr) {
    set_error("carlb_query_env: null output");
    return CARLB_ERR_INVALID;
  }
  memset(out, 0, sizeof(*out));
  switch (kind) {
    case KIND_CARTPOLE: fill_info<KIND_CARTPOLE>(out); return CARLB_OK;
    case KIND_PENDULUM: fill_info<KIND_PENDULUM>(out); return CARLB_OK;
    case KIND_ACROBOT: fill_info<KIND_ACROBOT>(out); return CARLB_OK;
    case KIND_MOUNTAINCAR: fill_info<KIND_MOUNTAINCAR>(out); return CARLB_OK;
    case KIND_MOUNTAINCAR_CONT: fill_info<KIND_MOUNTAINCAR_CONT>(out); return CARLB_OK;
    default: break;
  }
  if (is_brax(kind)) return brax_query(kind, out);
  set_error("carlb_query_env: unknown env kind %d", kind);
  return CARLB_ERR_INVALID;
}

int carlb_env_create(int kind, int n_envs, int precision, int device, int64_t global_offset, carlb_env_t** out) {
  if (out == nullptr) {
    set_error("carlb_env_create: null output");
    return CARLB_ERR_INVALID;
  }
  *out = nullptr;
  carlb_env_info_t info;
  if (carlb_query_env(kind, &info) != CARLB_OK) return CARLB_ERR_INVALID;
  if (n_envs <= 0) {
    set_error("carlb_env_create: n_envs must be positive, got %d", n_envs);
    return CARLB_ERR_INVALID;
  }
  if (precision != CARLB_F32 && precision != CARLB_F64) {
    set_error("carlb_env_create: unknown precision %d", precision);
    return CARLB_ERR_INVALID;
  }
  if (is_brax(kind) && precision != CARLB_F32) {
    set_error("carlb_env_create: Brax envs are float32 only (the reference's JAX pipeline is float32)");
    return CARLB_ERR_INVALID;
  }
  int n_dev = 0;
  CARLB_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    set_error("carlb_env_create: device %d out of range (%d visible)", device, n_dev);
    return CARLB_ERR_INVALID;
  }
  carlb_env* env = new (std::nothrow) carlb_env();
  if (env == nullptr) {
    set_error("carlb_env_create: out of host memory");
    return CARLB_ERR_STATE;
  }
  env->kind = kind;
  env->n = n_envs;
  env->precision = precision;
  env->device = device;
  env->max_steps = info.default_max_steps;
  env->autoreset = is_brax(kind) ? CARLB_AUTORESET_SAME_STEP : CARLB_AUTORESET_NONE;
  env->global_offset = (long long)global_offset;
  if (is_brax(kind)) {
    CARLB_CUDA_CHECK(cudaSetDevice(device));
    const int rc = brax_create(env);
    if (rc != CARLB_OK) {
      delete env;
      return rc;
    }
  }
  *out = env;
  return CARLB_OK;
}

int carlb_env_destroy(carlb_env_t* env) {
  if (env == nullptr) return CARLB_OK;
  if (env->gather != nullptr) gather_forget_env(env->gather, env);
  if (is_brax(env->kind)) brax_destroy(env);
  if (env->undo_block != nullptr) cudaFree(env->undo_block);
  if (env->bad_action_host != nullptr) cudaFreeHost(env->bad_action_host);
  if (env->part_counters != nullptr) cudaFree(env->part_counters);
  delete env;
  return CARLB_OK;
}

int carlb_env_bind(carlb_env_t* env, const carlb_buffers_t* b) {
  if (env == nullptr || b == nullptr) {
    set_error("carlb_env_bind: null argument");
    return CARLB_ERR_INVALID;
  }
  const bool need_sbt = env->kind == KIND_CARTPOLE;
  if (!b->state || !b->ctx || !b->elapsed || !b->rng || !b->obs || !b->reward || !b->terminated || !b->truncated ||
      (need_sbt && !b->sbt)) {
    set_error("carlb_env_bind: a required buffer pointer is null");
    return CARLB_ERR_INVALID;
  }
  if (is_brax(env->kind) && (!b->first_state || !b->first_obs)) {
    set_error("carlb_env_bind: Brax handles need first_state / first_obs buffers");
    return CARLB_ERR_INVALID;
  }
  if (((uintptr_t)b->state & 15u) || ((uintptr_t)b->obs & 15u) || ((uintptr_t)b->ctx & 7u) || ((uintptr_t)b->rng & 7u)) {
    set_error("carlb_env_bind: state/obs must be 16-byte aligned, ctx/rng 8-byte aligned");
    return CARLB_ERR_INVALID;
  }
  env->bufs = *b;
  env->bound = true;
  return CARLB_OK;
}

int carlb_env_configure(carlb_env_t* env, int max_episode_steps, int autoreset) {
  if (env == nullptr) {
    set_error("carlb_env_configure: null handle");
    return CARLB_ERR_INVALID;
  }
  if (autoreset != CARLB_AUTORESET_NONE && autoreset != CARLB_AUTORESET_SAME_STEP) {
    set_error("carlb_env_configure: unknown autoreset mode %d", autoreset);
    return CARLB_ERR_INVALID;
  }
  if (max_episode_steps > 0) env->max_steps = max_episode_steps;
  env->autoreset = autoreset;
  return CARLB_OK;
}

int carlb_env_seed(carlb_env_t* env, uint64_t seed, void* stream) {
  int rc = check_ready(env, "carlb_env_seed");
  if (rc != CARLB_OK) return rc;
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  if (is_brax(env->kind)) return brax_seed(env, seed, (cudaStream_t)stream);
  return classic_seed(env, seed, (cudaStream_t)stream);
}

int carlb_env_reset(carlb_env_t* env, const uint8_t* mask, void* stream) {
  int rc = check_ready(env, "carlb_env_reset");
  if (rc != CARLB_OK) return rc;
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  if (is_brax(env->kind)) return brax_reset(env, mask, (cudaStream_t)stream);
  return classic_reset(env, mask, (cudaStream_t)stream);
}

int carlb_env_step(carlb_env_t* env, const void* actions, int act_dtype, void* stream) {
  int rc = check_ready(env, "carlb_env_step");
  if (rc != CARLB_OK) return rc;
  if (actions == nullptr || !valid_act_dtype(env, act_dtype)) {
    set_error("carlb_env_step: null actions or action dtype %d not valid for env kind %d", act_dtype, env->kind);
    return CARLB_ERR_INVALID;
  }
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  if (is_brax(env->kind))
    return env->brax_arithmetic == CARLB_BRAX_FMA ? brax_step_fma(env, actions, act_dtype, (cudaStream_t)stream)
                                                  : brax_step(env, actions, act_dtype, (cudaStream_t)stream);
  return classic_step(env, actions, act_dtype, (cudaStream_t)stream);
}

int carlb_env_step_host(carlb_env_t* env, const void* actions_host, int act_dtype, float* obs_host, float* reward_host,
                        uint8_t* terminated_host, uint8_t* truncated_host, void* stream) {
  int rc = check_ready(env, "carlb_env_step_host");
  if (rc != CARLB_OK) return rc;
  if (actions_host == nullptr || !valid_act_dtype(env, act_dtype)) {
    set_error("carlb_env_step_host: null actions or action dtype %d not valid for env kind %d", act_dtype, env->kind);
    return CARLB_ERR_INVALID;
  }
  if (env->bufs.act_staging == nullptr) {
    set_error("carlb_env_step_host: act_staging buffer was not bound");
    return CARLB_ERR_STATE;
  }
  carlb_env_info_t info;
  carlb_query_env(env->kind, &info);
  cudaStream_t st = (cudaStream_t)stream;
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  const size_t esz = act_dtype == CARLB_ACT_I64 ? 8 : (act_dtype == CARLB_ACT_U8 ? 1 : 4);
  const size_t n = (size_t)env->n;
  // Zero-copy path (classic envs, all five host buffers page-locked): the kernel reads the actions
  // from and writes the results to mapped host memory itself -- no staging copy, no four
  // device->host copies, the PCIe traffic overlaps the compute; one stream sync ends the call.
  // CARLB_ZEROCOPY: 1 (default) as above; 2 = results zero-copy, actions by an async H2D copy into the
  // staging buffer (PCIe reads issued by the copy engine instead of by the SMs); 0 = staged copies.
  const char* zc_env = getenv("CARLB_ZEROCOPY");
  const int zero_copy = zc_env ? (zc_env[0] - '0') : 1;
  if (zero_copy > 0 && obs_host && reward_host && terminated_host && truncated_host &&
      is_mapped_host(actions_host) && is_mapped_host(obs_host) && is_mapped_host(reward_host) &&
      is_mapped_host(terminated_host) && is_mapped_host(truncated_host)) {
    HostMirrors hm{obs_host, reward_host, terminated_host, truncated_host};
    const void* act_src = actions_host;
    if (zero_copy == 2) {
      CARLB_CUDA_CHECK(cudaMemcpyAsync(env->bufs.act_staging, actions_host, n * esz * (size_t)info.act_dim,
                                       cudaMemcpyHostToDevice, st));
      act_src = env->bufs.act_staging;
    }
    rc = !is_brax(env->kind) ? classic_step(env, act_src, act_dtype, st, &hm)
         : env->brax_arithmetic == CARLB_BRAX_FMA ? brax_step_fma(env, act_src, act_dtype, st, &hm)
                                                  : brax_step(env, act_src, act_dtype, st, &hm);
    if (rc != CARLB_OK) return rc;
    CARLB_CUDA_CHECK(cudaStreamSynchronize(st));
    return CARLB_OK;
  }
  CARLB_CUDA_CHECK(cudaMemcpyAsync(env->bufs.act_staging, actions_host, n * esz * (size_t)info.act_dim,
                                   cudaMemcpyHostToDevice, st));
  rc = !is_brax(env->kind) ? classic_step(env, env->bufs.act_staging, act_dtype, st)
       : env->brax_arithmetic == CARLB_BRAX_FMA ? brax_step_fma(env, env->bufs.act_staging, act_dtype, st)
                                                : brax_step(env, env->bufs.act_staging, act_dtype, st);
  if (rc != CARLB_OK) return rc;
  // One device->host copy when the caller laid obs | reward | terminated | truncated out back to
  // back on both sides (the Python host layer does): 4 copies -> 1 (each costs ~8 us of latency).
  const unsigned char* d0 = reinterpret_cast<const unsigned char*>(env->bufs.obs);
  unsigned char* h0 = reinterpret_cast<unsigned char*>(obs_host);
  const size_t ob = n * info.obs_dim * sizeof(float), rb = n * sizeof(float);
  const bool packed = obs_host && reward_host && terminated_host && truncated_host &&
                      reinterpret_cast<const unsigned char*>(env->bufs.reward) == d0 + ob &&
                      env->bufs.terminated == d0 + ob + rb && env->bufs.truncated == d0 + ob + rb + n &&
                      reinterpret_cast<unsigned char*>(reward_host) == h0 + ob && terminated_host == h0 + ob + rb &&
                      truncated_host == h0 + ob + rb + n;
  if (packed) {
    CARLB_CUDA_CHECK(cudaMemcpyAsync(h0, d0, ob + rb + 2 * n, cudaMemcpyDeviceToHost, st));
  } else {
    if (obs_host)
      CARLB_CUDA_CHECK(cudaMemcpyAsync(obs_host, env->bufs.obs, ob, cudaMemcpyDeviceToHost, st));
    if (reward_host)
      CARLB_CUDA_CHECK(cudaMemcpyAsync(reward_host, env->bufs.reward, rb, cudaMemcpyDeviceToHost, st));
    if (terminated_host)
      CARLB_CUDA_CHECK(cudaMemcpyAsync(terminated_host, env->bufs.terminated, n, cudaMemcpyDeviceToHost, st));
    if (truncated_host)
      CARLB_CUDA_CHECK(cudaMemcpyAsync(truncated_host, env->bufs.truncated, n, cudaMemcpyDeviceToHost, st));
  }
  CARLB_CUDA_CHECK(cudaStreamSynchronize(st));
  return CARLB_OK;
}

// Undo log + report word of the checked step, allocated once per handle on first use.
static int ensure_step_check(carlb_env* env, int n_actions, StepCheck* chk) {
  carlb_env_info_t info;
  carlb_query_env(env->kind, &info);
  const size_t n = (size_t)env->n;
  const size_t word = env->precision == CARLB_F64 ? 8 : 4;
  const size_t state_b = ((n * info.state_words * word + 15) / 16) * 16, el_b = ((n * 4 + 15) / 16) * 16,
               flag_b = ((n + 15) / 16) * 16, rng_b = 2 * n * 8, obs_b = ((n * info.obs_dim * 4 + 15) / 16) * 16;
  if (env->undo_block == nullptr) {
    CARLB_CUDA_CHECK(cudaMalloc(&env->undo_block, state_b + el_b + 2 * flag_b + rng_b + obs_b + el_b + 2 * flag_b));
    CARLB_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&env->bad_action_host), 2 * CARLB_MAX_PARTS * sizeof(int),
                                   cudaHostAllocMapped));
    memset(env->bad_action_host, 0, 2 * CARLB_MAX_PARTS * sizeof(int));
    CARLB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&env->part_counters), CARLB_MAX_PARTS * sizeof(unsigned int)));
    CARLB_CUDA_CHECK(cudaMemset(env->part_counters, 0, CARLB_MAX_PARTS * sizeof(unsigned int)));
  }
  unsigned char* p = static_cast<unsigned char*>(env->undo_block);
  chk->first = 0;
  chk->count = env->n;
  chk->done_word = nullptr;
  chk->done_ticket = 0;
  chk->part_counter = env->part_counters;
  chk->n_actions = n_actions;
  chk->bad_action = env->bad_action_host;
  chk->undo_state = p;
  chk->undo_rng = reinterpret_cast<uint64_t*>(p + state_b);
  chk->undo_elapsed = reinterpret_cast<int32_t*>(p + state_b + rng_b);
  chk->undo_sbt = p + state_b + rng_b + el_b;
  chk->undo_rng_flag = p + state_b + rng_b + el_b + flag_b;
  unsigned char* q = p + state_b + rng_b + el_b + 2 * flag_b;
  chk->undo_obs = reinterpret_cast<float*>(q);
  chk->undo_reward = reinterpret_cast<float*>(q + obs_b);
  chk->undo_flags = q + obs_b + el_b;  // [2][n]: needs 2 * n bytes <= 2 * flag_b
  return CARLB_OK;
}

int carlb_env_step_host_checked(carlb_env_t* env, const void* actions_host, int act_dtype, int n_actions, float* obs_host,
                                float* reward_host, uint8_t* terminated_host, uint8_t* truncated_host, void* stream) {
  int rc = check_ready(env, "carlb_env_step_host_checked");
  if (rc != CARLB_OK) return rc;
  if (actions_host == nullptr || !valid_act_dtype(env, act_dtype)) {
    set_error("carlb_env_step_host_checked: null actions or action dtype %d not valid for env kind %d", act_dtype, env->kind);
    return CARLB_ERR_INVALID;
  }
  carlb_env_info_t info;
  carlb_query_env(env->kind, &info);
  static const bool zc_on = [] {
    const char* e = getenv("CARLB_ZEROCOPY");
    return e == nullptr || e[0] == '1';
  }();
  const bool discrete_check = n_actions > 0 && info.act_discrete != 0;
  bool zero_copy = zc_on && discrete_check && is_classic(env->kind) && obs_host && reward_host && terminated_host &&
                   truncated_host;
  if (zero_copy) {
    // the four result pointers are verified once (the host layer passes the same page-locked block every
    // step); the action pointer -- which a caller may change freely -- every time
    const void* res[4] = {obs_host, reward_host, terminated_host, truncated_host};
    for (int k = 0; k < 4 && zero_copy; ++k) {
      if (env->zc_verified[k] == res[k]) continue;
      if (is_mapped_host(res[k])) env->zc_verified[k] = res[k];
      else zero_copy = false;
    }
    zero_copy = zero_copy && is_mapped_host(actions_host);
  }
  if (!zero_copy) {  // staged path: range check on the host first, then the plain host-buffer step
    if (discrete_check) {
      rc = check_only(actions_host, (int64_t)env->n * info.act_dim, act_dtype, n_actions);
      if (rc != CARLB_OK) return rc;
    }
    return carlb_env_step_host(env, actions_host, act_dtype, obs_host, reward_host, terminated_host, truncated_host, stream);
  }
  // No gather attached: the kernel stores a completion word into mapped host memory and the call polls it
  // instead of synchronising the stream (CARLB_HOST_POLL=0 restores cudaStreamSynchronize for A/B runs).
  static const bool poll = [] {
    const char* e = getenv("CARLB_HOST_POLL");
    return e == nullptr || e[0] != '0';
  }();
  if (poll && env->gather == nullptr && !env->part_pending[0]) {
    rc = carlb_env_step_host_begin(env, 0, 1, actions_host, act_dtype, n_actions, obs_host, reward_host, terminated_host,
                                   truncated_host, stream);
    if (rc != CARLB_OK) return rc;
    return carlb_env_step_host_end(env, 0);
  }
  cudaStream_t st = (cudaStream_t)stream;
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  StepCheck chk;
  rc = ensure_step_check(env, n_actions, &chk);
  if (rc != CARLB_OK) return rc;
  HostMirrors hm{obs_host, reward_host, terminated_host, truncated_host};
  rc = classic_step_checked(env, actions_host, act_dtype, st, &hm, chk);
  if (rc != CARLB_OK) return rc;
  CARLB_CUDA_CHECK(cudaStreamSynchronize(st));
  const int bad = *reinterpret_cast<volatile int*>(env->bad_action_host);
  if (bad != 0) {  // roll every env back: the reference's env is untouched when `action_space.contains` fails
    *env->bad_action_host = 0;
    rc = classic_step_undo(env, st, chk, &hm);
    if (rc != CARLB_OK) return rc;
    CARLB_CUDA_CHECK(cudaStreamSynchronize(st));
    set_error("invalid action: values must lie in [0, %d) (env %d)", n_actions, bad - 1);
    return CARLB_ERR_INVALID;
  }
  return CARLB_OK;
}


// ---- split-batch host step: begin (enqueue one part, no sync) / end (poll the completion word)
static inline void part_range(int n, int part, int n_parts, int* first, int* count) {
  const long long lo = (long long)n * part / n_parts, hi = (long long)n * (part + 1) / n_parts;
  *first = (int)lo;
  *count = (int)(hi - lo);
}

int carlb_env_step_host_begin(carlb_env_t* env, int part, int n_parts, const void* actions_host, int act_dtype, int n_actions,
                              float* obs_host, float* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                              void* stream) {
  int rc = check_ready(env, "carlb_env_step_host_begin");
  if (rc != CARLB_OK) return rc;
  if (n_parts < 1 || n_parts > CARLB_MAX_PARTS || part < 0 || part >= n_parts || n_parts > env->n) {
    set_error("carlb_env_step_host_begin: part %d of %d (at most %d parts, each at least one env)", part, n_parts, CARLB_MAX_PARTS);
    return CARLB_ERR_INVALID;
  }
  if (!is_classic(env->kind) || env->gather != nullptr) {
    set_error("carlb_env_step_host_begin: classic-control handles without a fused gather only");
    return CARLB_ERR_INVALID;
  }
  if (actions_host == nullptr || !valid_act_dtype(env, act_dtype) || !obs_host || !reward_host || !terminated_host ||
      !truncated_host) {
    set_error("carlb_env_step_host_begin: null buffer or action dtype %d not valid for env kind %d", act_dtype, env->kind);
    return CARLB_ERR_INVALID;
  }
  if (env->part_pending[part]) {
    set_error("carlb_env_step_host_begin: part %d already has a step in flight (call carlb_env_step_host_end first)", part);
    return CARLB_ERR_STATE;
  }
  const void* res[4] = {obs_host, reward_host, terminated_host, truncated_host};
  for (int k = 0; k < 4; ++k) {
    if (env->zc_verified[k] == res[k]) continue;
    if (!is_mapped_host(res[k])) {
      set_error("carlb_env_step_host_begin: result buffers must be page-locked (mapped) host memory");
      return CARLB_ERR_INVALID;
    }
    env->zc_verified[k] = res[k];
  }
  if (!is_mapped_host(actions_host)) {
    set_error("carlb_env_step_host_begin: the action array must be page-locked (mapped) host memory");
    return CARLB_ERR_INVALID;
  }
  carlb_env_info_t info;
  carlb_query_env(env->kind, &info);
  cudaStream_t st = (cudaStream_t)stream;
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  StepCheck chk;
  rc = ensure_step_check(env, info.act_discrete ? n_actions : 0, &chk);
  if (rc != CARLB_OK) return rc;
  part_range(env->n, part, n_parts, &chk.first, &chk.count);
  chk.bad_action = env->bad_action_host + part;
  chk.done_word = reinterpret_cast<unsigned int*>(env->bad_action_host) + CARLB_MAX_PARTS + part;
  chk.done_ticket = ++env->part_ticket[part];
  if (chk.done_ticket == 0) chk.done_ticket = ++env->part_ticket[part];  // 0 is the initial value of the word
  chk.part_counter = env->part_counters + part;
  const size_t esz = act_dtype == CARLB_ACT_I64 ? 8 : (act_dtype == CARLB_ACT_U8 ? 1 : 4);
  const size_t arow = esz * (size_t)info.act_dim;
  // (Moving the data with the copy engines instead -- actions host -> device staging, four result slices device ->
  // host, a one-thread kernel storing the completion word behind them -- was measured and dropped: every small
  // cudaMemcpyAsync costs microseconds of fixed latency, 70.9 vs 36.6 us per two-part step, profiles/r02j_e2e_async_probe.json.)
  HostMirrors hm{obs_host, reward_host, terminated_host, truncated_host};
  // the kernel indexes the action array with the env's index in the handle: shift the part's pointer back by
  // `first` elements (only the part's own rows are ever dereferenced)
  const unsigned char* act_base = static_cast<const unsigned char*>(actions_host) - (size_t)chk.first * arow;
  rc = classic_step_checked(env, act_base, act_dtype, st, &hm, chk);
  if (rc != CARLB_OK) return rc;
  env->part_pending[part] = true;
  env->part_first[part] = chk.first;
  env->part_count[part] = chk.count;
  env->part_n_actions[part] = chk.n_actions;
  env->part_stream[part] = st;
  return CARLB_OK;
}

int carlb_env_step_host_end(carlb_env_t* env, int part) {
  int rc = check_ready(env, "carlb_env_step_host_end");
  if (rc != CARLB_OK) return rc;
  if (part < 0 || part >= CARLB_MAX_PARTS || !env->part_pending[part]) {
    set_error("carlb_env_step_host_end: part %d has no step in flight", part);
    return CARLB_ERR_STATE;
  }
  volatile unsigned int* word = reinterpret_cast<volatile unsigned int*>(env->bad_action_host) + CARLB_MAX_PARTS + part;
  const unsigned int ticket = env->part_ticket[part];
  unsigned long long spins = 0;
  while (*word != ticket) {
    __builtin_ia32_pause();
    if ((++spins & 0xFFFFFull) == 0) {  // every ~1M polls: has the stream failed?
      const cudaError_t q = cudaStreamQuery(env->part_stream[part]);
      if (q != cudaSuccess && q != cudaErrorNotReady) {
        env->part_pending[part] = false;
        set_error("carlb_env_step_host_end: the step of part %d failed: %s", part, cudaGetErrorString(q));
        return CARLB_ERR_CUDA;
      }
    }
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  env->part_pending[part] = false;
  const int bad = *reinterpret_cast<volatile int*>(env->bad_action_host + part);
  if (bad != 0) {  // roll the part back: the reference's env is untouched when `action_space.contains` fails
    env->bad_action_host[part] = 0;
    StepCheck chk;
    rc = ensure_step_check(env, env->part_n_actions[part], &chk);
    if (rc != CARLB_OK) return rc;
    chk.first = env->part_first[part];
    chk.count = env->part_count[part];
    CARLB_CUDA_CHECK(cudaSetDevice(env->device));
    const HostMirrors hm{const_cast<float*>(static_cast<const float*>(env->zc_verified[0])),
                         const_cast<float*>(static_cast<const float*>(env->zc_verified[1])),
                         const_cast<uint8_t*>(static_cast<const uint8_t*>(env->zc_verified[2])),
                         const_cast<uint8_t*>(static_cast<const uint8_t*>(env->zc_verified[3]))};
    rc = classic_step_undo(env, env->part_stream[part], chk, &hm);
    if (rc != CARLB_OK) return rc;
    CARLB_CUDA_CHECK(cudaStreamSynchronize(env->part_stream[part]));
    set_error("invalid action: values must lie in [0, %d) (env %d)", env->part_n_actions[part], bad - 1);
    return CARLB_ERR_INVALID;
  }
  return CARLB_OK;
}

int carlb_stage_actions(void* dst_pinned, const void* src, int64_t count, int act_dtype, int n_actions) {
  if (src == nullptr || count < 0) {
    set_error("carlb_stage_actions: null source or negative count");
    return CARLB_ERR_INVALID;
  }
  if (dst_pinned == nullptr || dst_pinned == src) return check_only(src, count, act_dtype, n_actions);
  bool ok = true;
  switch (act_dtype) {
    case CARLB_ACT_I32: ok = stage_checked((int32_t*)dst_pinned, (const int32_t*)src, count, (uint64_t)n_actions); break;
    case CARLB_ACT_I64: ok = stage_checked((int64_t*)dst_pinned, (const int64_t*)src, count, (uint64_t)n_actions); break;
    case CARLB_ACT_U8: ok = stage_checked((uint8_t*)dst_pinned, (const uint8_t*)src, count, (uint64_t)n_actions); break;
    case CARLB_ACT_F32: memcpy(dst_pinned, src, (size_t)count * sizeof(float)); break;
    default: set_error("carlb_stage_actions: unknown action dtype %d", act_dtype); return CARLB_ERR_INVALID;
  }
  if (!ok && n_actions > 0) {
    set_error("invalid action: values must lie in [0, %d)", n_actions);
    return CARLB_ERR_INVALID;
  }
  return CARLB_OK;
}

int carlb_env_rollout(carlb_env_t* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                      int act_dtype, const carlb_traj_t* traj, void* stream) {
  int rc = check_ready(env, "carlb_env_rollout");
  if (rc != CARLB_OK) return rc;
  if (n_steps < 0) {
    set_error("carlb_env_rollout: negative n_steps");
    return CARLB_ERR_INVALID;
  }
  if (actions != nullptr && !valid_act_dtype(env, act_dtype)) {
    set_error("carlb_env_rollout: action dtype %d not valid for env kind %d", act_dtype, env->kind);
    return CARLB_ERR_INVALID;
  }
  if (n_steps == 0) return CARLB_OK;  // nothing to do (and no observation produced: the fused gather must not count it)
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  if (is_brax(env->kind))
    return env->brax_arithmetic == CARLB_BRAX_FMA
               ? brax_rollout_fma(env, n_steps, policy_seed, step_base, actions, act_dtype, traj, (cudaStream_t)stream)
               : brax_rollout(env, n_steps, policy_seed, step_base, actions, act_dtype, traj, (cudaStream_t)stream);
  return classic_rollout(env, n_steps, policy_seed, step_base, actions, act_dtype, traj, (cudaStream_t)stream);
}

int carlb_mixed_step(carlb_env_t* const* envs, const void* const* actions, const int* act_dtypes, int n_handles,
                     void* stream) {
  if (envs == nullptr || actions == nullptr || act_dtypes == nullptr || n_handles <= 0 || n_handles > CARLB_MAX_MIXED) {
    set_error("carlb_mixed_step: bad arguments (n_handles=%d, max %d)", n_handles, CARLB_MAX_MIXED);
    return CARLB_ERR_INVALID;
  }
  for (int k = 0; k < n_handles; ++k) {
    int rc = check_ready(envs[k], "carlb_mixed_step");
    if (rc != CARLB_OK) return rc;
    if (!is_classic(envs[k]->kind)) {
      set_error("carlb_mixed_step: handle %d is not a classic-control shard", k);
      return CARLB_ERR_INVALID;
    }
    if (envs[k]->device != envs[0]->device) {
      set_error("carlb_mixed_step: all shards of one launch must live on one device");
      return CARLB_ERR_INVALID;
    }
    if (actions[k] == nullptr || !valid_act_dtype(envs[k], act_dtypes[k])) {
      set_error("carlb_mixed_step: bad actions for handle %d", k);
      return CARLB_ERR_INVALID;
    }
  }
  CARLB_CUDA_CHECK(cudaSetDevice(envs[0]->device));
  return classic_mixed_step(envs, actions, act_dtypes, n_handles, (cudaStream_t)stream);
}

int carlb_brax_set_system(carlb_env_t* env, const float* table, int n_floats, int stock_contact) {
  if (env == nullptr || table == nullptr || !is_brax(env->kind)) {
    set_error("carlb_brax_set_system: needs a Brax handle and a table");
    return CARLB_ERR_INVALID;
  }
  return brax_set_system(env, table, n_floats, stock_contact);
}

int carlb_brax_set_arithmetic(carlb_env_t* env, int arithmetic) {
  if (env == nullptr || !is_brax(env->kind) || (arithmetic != CARLB_BRAX_STRICT && arithmetic != CARLB_BRAX_FMA)) {
    set_error("carlb_brax_set_arithmetic: needs a Brax handle and CARLB_BRAX_STRICT / CARLB_BRAX_FMA");
    return CARLB_ERR_INVALID;
  }
  env->brax_arithmetic = arithmetic;
  return CARLB_OK;
}

int carlb_brax_set_reset_rng(carlb_env_t* env, int mode, int64_t n_global) {
  if (env == nullptr || !is_brax(env->kind) || (mode != CARLB_RESET_PHILOX && mode != CARLB_RESET_JAX) || n_global < 1) {
    set_error("carlb_brax_set_reset_rng: needs a Brax handle, CARLB_RESET_PHILOX / CARLB_RESET_JAX and n_global >= 1");
    return CARLB_ERR_INVALID;
  }
  return brax_set_reset_rng(env, mode, (long long)n_global);
}

int carlb_brax_reset_from_q(carlb_env_t* env, const uint8_t* mask, const float* q, const float* qd, void* stream) {
  int rc = check_ready(env, "carlb_brax_reset_from_q");
  if (rc != CARLB_OK) return rc;
  if (!is_brax(env->kind) || q == nullptr || qd == nullptr) {
    set_error("carlb_brax_reset_from_q: needs a Brax handle and q / qd");
    return CARLB_ERR_INVALID;
  }
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  return brax_reset_from(env, mask, q, qd, (cudaStream_t)stream);
}

int carlb_brax_goal_step(carlb_env_t* env, int idx0, int idx1, double dt, double* position, const double* goal,
                         const double* radius, double* reward, uint8_t* success, void* stream) {
  int rc = check_ready(env, "carlb_brax_goal_step");
  if (rc != CARLB_OK) return rc;
  carlb_env_info_t info;
  if (!is_brax(env->kind) || position == nullptr || goal == nullptr || radius == nullptr || reward == nullptr ||
      success == nullptr) {
    set_error("carlb_brax_goal_step: needs a Brax handle and position / goal / radius / reward / success buffers");
    return CARLB_ERR_INVALID;
  }
  brax_query(env->kind, &info);
  if (idx0 < 0 || idx1 < 0 || idx0 >= info.obs_dim || idx1 >= info.obs_dim) {
    set_error("carlb_brax_goal_step: observation indices (%d, %d) outside [0, %d)", idx0, idx1, info.obs_dim);
    return CARLB_ERR_INVALID;
  }
  CARLB_CUDA_CHECK(cudaSetDevice(env->device));
  return brax_goal_step(env, idx0, idx1, dt, position, goal, radius, reward, success, (cudaStream_t)stream);
}

}  // extern "C"
