// Shared device-side declarations of libcarlb (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/carlb.h"
#include "physics_classic.h"

namespace carlb {

// ------------------------------------------------------------------------------------------------
// Fused cross-GPU observation gather, device side (host side: gather.cu).
//
// Every rank owns one symmetric buffer  obs[kGatherSlots][n_global][D] | flags[MAX_PEERS] | ctrl[8]
// mapped into all ranks. A "push" stores this rank's rows into the SAME slot of every rank's buffer
// and then publishes the push count into its flag word on every rank. Slot and flag value come from
// a counter in DEVICE memory (ctrl[0] = pushes published so far), not from kernel parameters, so a
// sequence of obs-producing launches can be captured into a CUDA graph and replayed.
//
//   GATHER_IMMEDIATE  the rows this launch computes are stored at its end (step / reset / rollout /
//                     Brax kernels), the last CTA publishes, then (wait_lag >= 0) spins until every
//                     rank has published push  seq + 1 - wait_lag.
//   GATHER_DEFERRED   (classic step / rollout) a dedicated publisher warp per CTA pushes the rows the
//                     PREVIOUS launch left in seg.obs while the other warps compute: the NVLink
//                     transfer and its system-scope fence overlap the physics instead of draining at
//                     the end of the grid.
constexpr int kGatherSlots = 4;
enum { GATHER_IMMEDIATE = 0, GATHER_DEFERRED = 1 };
enum { GCTRL_SEQ = 0, GCTRL_PUSH_COUNTER = 1 };

struct GatherDev {
  int n_peers;                                // world size (0: no gather attached)
  int mode;                                   // GATHER_IMMEDIATE / GATHER_DEFERRED
  int wait_lag;                               // < 0: no in-kernel wait
  int debug;                                  // CARLB_GATHER_DEBUG bit mask (A/B measurements only; breaks correctness)
  unsigned long long slot_floats;             // floats per slot
  float* peer_base[CARLB_MAX_PEERS];          // slot 0 of rank r's buffer (mapped here)
  unsigned int* peer_flags[CARLB_MAX_PEERS];  // MY flag word in rank r's buffer
  float* mc_base;                             // multicast alias of slot 0 (one multimem.st reaches every rank) or null
  unsigned int* mc_flag;                      // multicast alias of my flag word, or null
  const unsigned int* my_flags;               // flags[world] of the local buffer (written by the peers)
  unsigned int* ctrl;                         // local control words (GCTRL_*)
};

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Push sequence number of this launch: every thread reads the device-side counter itself, first thing in the
// kernel (no barrier: the load overlaps the thread's other prologue loads). All threads of the grid see the same
// value because the counter only moves when the LAST CTA publishes, i.e. after every CTA has arrived, and a CTA
// arrives only after all of its threads have passed this read.
// (No "memory" clobber on this load, on the row stores or on the compute warps' barrier arrival: they are
// `asm volatile`, so the compiler keeps them in order among themselves, but it stays free to hoist the thread's
// own prologue loads -- state, context rows -- above them; otherwise every compute thread would first wait an L2
// round trip for the counter and only then start loading its env.)
__device__ __forceinline__ unsigned int gather_begin(const GatherDev& g) {
  if (g.n_peers <= 0) return 0u;
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(g.ctrl + GCTRL_SEQ));
  return v;
}

// Store one obs row (D floats, 16-byte aligned when D % 4 == 0) into slot `seq` of every rank.
template <int D>
__device__ __forceinline__ void gather_store_row(const GatherDev& g, unsigned int seq, size_t global_row, const float* o) {
  const size_t off = (size_t)(seq % kGatherSlots) * g.slot_floats + global_row * D;
  if (g.debug & 4) return;
  if (g.mc_base != nullptr) {
    float* dst = g.mc_base + off;
    if (D % 4 == 0) {
#pragma unroll
      for (int k = 0; k < D; k += 4)
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "f"(o[k]), "f"(o[k + 1]),
                     "f"(o[k + 2]), "f"(o[k + 3]));
    } else if (D % 2 == 0) {
#pragma unroll
      for (int k = 0; k < D; k += 2)
        asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(dst + k), "f"(o[k]), "f"(o[k + 1]));
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(dst + k), "f"(o[k]));
    }
    return;
  }
  for (int r = 0; r < g.n_peers; ++r) {
    float* dst = g.peer_base[r] + off;
    if (D % 4 == 0) {
#pragma unroll
      for (int k = 0; k < D; k += 4)
        asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "f"(o[k]), "f"(o[k + 1]), "f"(o[k + 2]),
                     "f"(o[k + 3]));
    } else if (D % 2 == 0) {
#pragma unroll
      for (int k = 0; k < D; k += 2) asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(dst + k), "f"(o[k]), "f"(o[k + 1]));
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) asm volatile("st.global.f32 [%0], %1;" ::"l"(dst + k), "f"(o[k]));
    }
  }
}
// One float of an obs row (kernels whose lanes own single elements: Brax).
__device__ __forceinline__ void gather_store_elem(const GatherDev& g, unsigned int seq, size_t global_elem, float v) {
  const size_t off = (size_t)(seq % kGatherSlots) * g.slot_floats + global_elem;
  if (g.mc_base != nullptr) {
    asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(g.mc_base + off), "f"(v));
    return;
  }
  for (int r = 0; r < g.n_peers; ++r) asm volatile("st.global.f32 [%0], %1;" ::"l"(g.peer_base[r] + off), "f"(v));
}

// Called by ONE thread per CTA after the CTA's pushed rows are ordered before it (barrier). The CTA
// arrives with a DEVICE-scope release (fence + counter); only the last CTA to arrive pays the
// system-scope fence and publishes push `seq` (flag value seq + 1) to every rank, then advances the
// device-side sequence counter. Cumulativity makes that single system fence cover every CTA's rows: their
// stores happen-before the arrivals the last CTA has observed (release / acquire at gpu scope), which
// happen-before its fence.sys and the release store of the flag. One MEMBAR.SYS per launch instead of one
// per CTA matters: system-scope fences of different SMs serialise (r02b: 1 024 of them cost ~12 us per launch).
// Returns true in the thread that published.
__device__ __forceinline__ bool gather_publish(const GatherDev& g, unsigned int seq, unsigned int n_ctas) {
  if (!(g.debug & 1)) __threadfence();
  const unsigned int prev = atomicAdd(g.ctrl + GCTRL_PUSH_COUNTER, 1u);
  if (prev != n_ctas - 1) return false;
  __threadfence();  // acquire side of the CTA arrivals
  g.ctrl[GCTRL_PUSH_COUNTER] = 0;
  // release pattern: ONE system-scope fence, then relaxed flag stores (a st.release per peer would repeat the
  // MEMBAR.SYS for every rank)
  if (!(g.debug & 8)) __threadfence_system();
  if (g.mc_flag != nullptr) {
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(g.mc_flag), "r"(seq + 1u) : "memory");
  } else {
    for (int r = 0; r < g.n_peers; ++r)
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(g.peer_flags[r]), "r"(seq + 1u) : "memory");
  }
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(g.ctrl + GCTRL_SEQ), "r"(seq + 1u) : "memory");
  return true;
}

// Spin (one thread) until every rank has published push `seq - wait_lag`.
__device__ __forceinline__ void gather_wait_all(const GatherDev& g, unsigned int seq) {
  if (g.wait_lag < 0 || (g.debug & 2)) return;
  const int target = (int)(seq + 1u) - g.wait_lag;
  if (target <= 0) return;
  for (int r = 0; r < g.n_peers; ++r)
    while ((int)(ld_acquire_sys_u32(g.my_flags + r) - (unsigned int)target) < 0) {}
}

// Epilogue of an IMMEDIATE push: every thread of every CTA calls it once, from non-divergent code,
// after its row stores. The thread that publishes also performs the in-kernel wait, so the completion
// of the grid implies that the gathered tensor of push `seq - wait_lag` is complete on this rank.
__device__ __forceinline__ void gather_epilogue_immediate(const GatherDev& g, unsigned int seq) {
  if (g.n_peers <= 0) return;
  __syncthreads();  // the CTA's row stores happen-before thread 0's fence (cumulativity through the barrier)
  if (threadIdx.x == 0 && gather_publish(g, seq, gridDim.x)) gather_wait_all(g, seq);
}

// Named barrier of the DEFERRED push (barrier 0 is __syncthreads). `count` must be a multiple of 32 and every
// thread of a participating warp must execute the instruction.
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// the compute warps' side: ordered after the thread's (asm volatile) row stores, no compiler memory barrier
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count)); }
constexpr int kBarPushed = 1;   // step kernel: compute warps arrive after storing their rows; the publisher warp syncs on it
constexpr int kBarObsRead = 2;  // rollout kernel: the publisher warp arrives after reading the old rows; compute warps sync before overwriting them

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
  // fused cross-GPU observation gather of this launch (gth.n_peers == 0: none)
  GatherDev gth;
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
  int first;               // this launch steps the envs [first, first + count) of the handle (a "part")
  int count;
  int n_actions;           // valid discrete actions are [0, n_actions); <= 0: no check
  unsigned int* done_word;     // mapped host word: the last CTA stores done_ticket after a system fence (or null)
  unsigned int done_ticket;
  unsigned int* part_counter;  // device: CTA arrival counter of this part
  int* bad_action;         // mapped host word: 1 + index of an env whose action is invalid (0: none)
  void* undo_state;        // T[n][S] state before this step
  int32_t* undo_elapsed;   // [n]
  uint8_t* undo_sbt;       // [n]
  uint64_t* undo_rng;      // [2][n] PCG64 state words before this step (valid where undo_rng_flag)
  uint8_t* undo_rng_flag;  // [n]
  // what the step RETURNS is logged too, so that after a roll-back the device result buffers and the caller's host
  // arrays show the previous step again (`state_dict()` right after a rejected step must be consistent)
  float* undo_obs;         // [n][D]
  float* undo_reward;      // [n]
  uint8_t* undo_flags;     // [2][n] terminated, truncated
};

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
