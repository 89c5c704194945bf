// Internal host-side declarations shared by abi.cu / classic.cu / brax.cu.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "common.cuh"

struct carlb_gather;

struct carlb_env {
  int kind = 0;
  int n = 0;
  int precision = CARLB_F32;
  int device = 0;
  int max_steps = 0;
  int autoreset = CARLB_AUTORESET_NONE;
  long long global_offset = 0;
  bool bound = false;
  carlb_buffers_t bufs{};
  int brax_arithmetic = CARLB_BRAX_STRICT;  // which build of the Brax step / rollout kernels this handle runs
  void* brax_sys = nullptr;  // BraxHandle: device copy of the per-handle Brax system table
  carlb_gather* gather = nullptr;  // fused cross-GPU obs gather (gather.cu), or null
  // host-buffer step with in-kernel action validation (carlb_env_step_host_checked): undo log (one device
  // block, allocated on first use -- never per step) and the mapped host word the kernel reports into
  void* undo_block = nullptr;
  unsigned int* part_counters = nullptr;  // device: CTA arrival counters of the completion words
  int* bad_action_host = nullptr;   // mapped host block: [CARLB_MAX_PARTS] bad-action words, then [CARLB_MAX_PARTS] completion words
  const void* zc_verified[4] = {};  // result pointers already verified as mapped page-locked memory
  // split-batch host step (carlb_env_step_host_begin / _end): per part, the ticket the kernel will store and
  // what is needed to roll the part back
  unsigned int part_ticket[CARLB_MAX_PARTS] = {};
  bool part_pending[CARLB_MAX_PARTS] = {};
  int part_first[CARLB_MAX_PARTS] = {}, part_count[CARLB_MAX_PARTS] = {}, part_n_actions[CARLB_MAX_PARTS] = {};
  cudaStream_t part_stream[CARLB_MAX_PARTS] = {};
};

namespace carlb {

extern std::atomic<long long> g_launches;
void set_error(const char* fmt, ...);
Segment make_segment(const carlb_env* env, int act_dtype);

// gather.cu
void gather_forget_env(carlb_gather* g, carlb_env* env);
// kinds of obs-producing launches, as the gather's bookkeeping sees them
enum GatherLaunch {
  GL_RESET = 0,           // immediate push of the rows the launch computes, waits for every rank's push of this launch
  GL_CLASSIC_STEP = 1,    // classic step / rollout: the push may be deferred to a publisher warp (pipelined mode)
  GL_BRAX = 2,            // Brax step / rollout: immediate push; pipelined mode waits one push behind
  GL_IMMEDIATE_ONLY = 3,  // checked host-buffer step: as GL_RESET
};
void gather_fill(carlb_gather* g, GatherDev* out, int launch_kind);

// classic.cu
int classic_seed(const carlb_env* env, uint64_t seed, cudaStream_t st);
int classic_reset(const carlb_env* env, const uint8_t* mask, cudaStream_t st);
struct HostMirrors {  // mapped (page-locked) host result buffers of one zero-copy step
  float* obs; float* reward; uint8_t* terminated; uint8_t* truncated;
};
int classic_step(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm = nullptr);
int classic_step_checked(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm,
                         const StepCheck& chk);
int classic_step_undo(const carlb_env* env, cudaStream_t st, const StepCheck& chk, const HostMirrors* hm = nullptr);
int classic_rollout(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                    int act_dtype, const carlb_traj_t* traj, cudaStream_t st);
int classic_mixed_step(carlb_env* const* envs, const void* const* actions, const int* act_dtypes, int n_handles,
                       cudaStream_t st);

// brax.cu
int brax_query(int kind, carlb_env_info_t* out);
int brax_create(carlb_env* env);
void brax_destroy(carlb_env* env);
int brax_seed(const carlb_env* env, uint64_t seed, cudaStream_t st);
int brax_reset(const carlb_env* env, const uint8_t* mask, cudaStream_t st);
int brax_step(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm = nullptr);
int brax_rollout(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                 int act_dtype, const carlb_traj_t* traj, cudaStream_t st);
// brax_fma.cu: the FMA-contracted build of the step / rollout kernels
int brax_step_fma(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm = nullptr);
int brax_rollout_fma(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                     int act_dtype, const carlb_traj_t* traj, cudaStream_t st);
int brax_set_system(carlb_env* env, const float* table, int n_floats, int stock_contact);
int brax_reset_from(const carlb_env* env, const uint8_t* mask, const float* q, const float* qd, cudaStream_t st);
int brax_set_reset_rng(carlb_env* env, int mode, long long n_global);
int brax_goal_step(const carlb_env* env, int idx0, int idx1, double dt, double* position, const double* goal,
                   const double* radius, double* reward, uint8_t* success, cudaStream_t st);

#define CARLB_CUDA_CHECK(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      carlb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CARLB_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

}  // namespace carlb
