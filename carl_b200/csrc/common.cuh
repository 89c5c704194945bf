// Shared device-side declarations of libcarlb (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/carlb.h"
#include "physics_classic.h"

namespace carlb {

// One homogeneous shard of env instances resident on this GPU. All pointers are device
// pointers into caller-owned buffers (torch tensors); the library never allocates per step.
struct Segment {
  int kind;
  int n;               // env instances in this shard
  int max_steps;       // TimeLimit / EpisodeWrapper length (<=0: none)
  int autoreset;       // 0: none (reference classic-control behaviour), 1: same-step auto-reset
  int act_dtype;       // CARLB_ACT_*
  long long global_offset;  // global env id of local env 0 (keys RNG streams; sharding-invariant)
  void* state;         // T[n][S]
  const void* ctx;     // T[P][n]
  int32_t* elapsed;    // [n]
  uint8_t* sbt;        // [n]  CartPole steps_beyond_terminated flag
  uint64_t* rng;       // [4][n] PCG64 (state_hi, state_lo, inc_hi, inc_lo)
  float* obs;          // [n][D]
  float* reward;       // [n]
  uint8_t* terminated; // [n]
  uint8_t* truncated;  // [n]
  float* final_obs;    // [n][D] or null
  // fused cross-GPU observation gather: obs rows are additionally stored at
  // peer_obs[r] + (global_offset + i) * D for every peer r (P2P-mapped symmetric buffers)
  int n_peers;
  float* peer_obs[CARLB_MAX_PEERS];
  unsigned int* peer_flags[CARLB_MAX_PEERS];  // fused gather: my completion word in every rank's buffer
  unsigned int signal_value;                  // value published when this launch's rows are stored
  unsigned int* block_counter;                // local CTA arrival counter (last CTA publishes)
  // zero-copy host mirrors (carlb_env_step_host with page-locked buffers): results are ALSO stored
  // straight into mapped host memory by the kernel (posted PCIe writes overlap the compute) instead
  // of four device->host copies after it; null otherwise
  float* host_obs;
  float* host_reward;
  uint8_t* host_terminated;
  uint8_t* host_truncated;
};

// In-kernel action validation + undo log of the host-buffer step (carlb_env_step_host_checked): the
// step kernel itself checks every discrete action against [0, n_actions) -- the `action_space.contains`
// assert of the gymnasium envs -- instead of a host pass over the action array before the launch, and
// logs what it overwrites so that a step with an invalid action can be rolled back (the reference's
// env is untouched when the assert fires).
struct StepCheck {
  int n_actions;           // valid discrete actions are [0, n_actions)
  int* bad_action;         // mapped host word: 1 + index of an env whose action is invalid (0: none)
  void* undo_state;        // T[n][S] state before this step
  int32_t* undo_elapsed;   // [n]
  uint8_t* undo_sbt;       // [n]
  uint64_t* undo_rng;      // [2][n] PCG64 state words before this step (valid where undo_rng_flag)
  uint8_t* undo_rng_flag;  // [n]
};

// Fused-gather epilogue: every thread of every CTA calls this at the end of an obs-producing
// kernel. Stores to peer memory are fenced at system scope, the last CTA to arrive publishes the
// launch number into every rank's flag word.
__device__ __forceinline__ void peer_signal_epilogue(int n_peers, unsigned int* const* peer_flags, unsigned int signal_value,
                                                     unsigned int* block_counter) {
  if (n_peers <= 0 || block_counter == nullptr) return;
  __syncthreads();  // the CTA's peer stores happen-before thread 0's fence (cumulativity through the barrier)
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(block_counter, 1u);
    if (prev == gridDim.x - 1) {
      *block_counter = 0;
      __threadfence_system();
      for (int r = 0; r < n_peers; ++r)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[r]), "r"(signal_value) : "memory");
    }
  }
}

__device__ __forceinline__ Action load_action(const void* actions, int act_dtype, long long i) {
  Action a;
  a.i = 0;
  a.f = 0.0f;
  switch (act_dtype) {
    case CARLB_ACT_I32: a.i = static_cast<const int32_t*>(actions)[i]; a.f = (float)a.i; break;
    case CARLB_ACT_I64: a.i = (int)static_cast<const long long*>(actions)[i]; a.f = (float)a.i; break;
    case CARLB_ACT_U8: a.i = (int)static_cast<const uint8_t*>(actions)[i]; a.f = (float)a.i; break;
    default: a.f = static_cast<const float*>(actions)[i]; a.i = (int)a.f; break;
  }
  return a;
}

// 128-bit (or narrower) vector load / store of one env's S-word state row.
template <typename T, int S> struct StateIO;
template <> struct StateIO<float, 4> {
  static __device__ __forceinline__ void load(const void* base, int i, float s[4]) {
    const float4 v = static_cast<const float4*>(base)[i];
    s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
  }
  static __device__ __forceinline__ void store(void* base, int i, const float s[4]) {
    static_cast<float4*>(base)[i] = make_float4(s[0], s[1], s[2], s[3]);
  }
};
template <> struct StateIO<float, 2> {
  static __device__ __forceinline__ void load(const void* base, int i, float s[2]) {
    const float2 v = static_cast<const float2*>(base)[i];
    s[0] = v.x; s[1] = v.y;
  }
  static __device__ __forceinline__ void store(void* base, int i, const float s[2]) {
    static_cast<float2*>(base)[i] = make_float2(s[0], s[1]);
  }
};
template <> struct StateIO<double, 4> {
  static __device__ __forceinline__ void load(const void* base, int i, double s[4]) {
    const double2 a = static_cast<const double2*>(base)[2 * i], b = static_cast<const double2*>(base)[2 * i + 1];
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
  }
  static __device__ __forceinline__ void store(void* base, int i, const double s[4]) {
    static_cast<double2*>(base)[2 * i] = make_double2(s[0], s[1]);
    static_cast<double2*>(base)[2 * i + 1] = make_double2(s[2], s[3]);
  }
};
template <> struct StateIO<double, 2> {
  static __device__ __forceinline__ void load(const void* base, int i, double s[2]) {
    const double2 a = static_cast<const double2*>(base)[i];
    s[0] = a.x; s[1] = a.y;
  }
  static __device__ __forceinline__ void store(void* base, int i, const double s[2]) {
    static_cast<double2*>(base)[i] = make_double2(s[0], s[1]);
  }
};

__device__ __forceinline__ Pcg64 load_rng(const uint64_t* rng, int n, int i) {
  Pcg64 g;
  g.state_hi = rng[i]; g.state_lo = rng[(size_t)n + i]; g.inc_hi = rng[2 * (size_t)n + i]; g.inc_lo = rng[3 * (size_t)n + i];
  return g;
}
__device__ __forceinline__ void store_rng_state(uint64_t* rng, int n, int i, const Pcg64& g) {
  rng[i] = g.state_hi; rng[(size_t)n + i] = g.state_lo;  // the increment never changes
}

}  // namespace carlb
