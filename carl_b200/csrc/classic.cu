// Classic-control kernels (sm_100a): one env instance per thread, state rows as 128-bit vector
// loads, per-env context SoA rows (coalesced 32-/64-bit loads), everything else in registers.
//
// These kernels are HBM/L2-latency bound integer+fp32 work (a 4-float state and ~20-300 flops per
// env-step): no tensor cores, no shared memory. Single-step launches implement the reference's
// `CARLEnv.step` contract; the fused rollout keeps the state in registers across K steps and
// streams the trajectory to HBM.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "engine.h"

namespace carlb {

constexpr int kBlock = 128;

template <int D> __device__ __forceinline__ void store_obs(float* base, size_t row, const float* o) {
  float* dst = base + row * D;
  if (D == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
  } else if (D == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
  } else if (D == 6) {
    float2* d2 = reinterpret_cast<float2*>(dst);
    d2[0] = make_float2(o[0], o[1]); d2[1] = make_float2(o[2], o[3]); d2[2] = make_float2(o[4], o[5]);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) dst[k] = o[k];
  }
}

template <int KIND, typename T>
__device__ __forceinline__ void load_rows(const Segment& seg, int i, int r0, int r1, T* p) {
  const T* ctx = static_cast<const T*>(seg.ctx);
#pragma unroll
  for (int r = 0; r < Traits<KIND>::P; ++r)
    if (r >= r0 && r < r1) p[r] = __ldg(ctx + (size_t)r * seg.n + i);
}

// ------------------------------------------------------------------------------ seed
__global__ void __launch_bounds__(kBlock) seed_kernel(uint64_t* rng, int n, uint64_t seed, long long global_offset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Pcg64 g;
  pcg64_seed_from_int(g, seed + (uint64_t)global_offset + (uint64_t)i);
  rng[i] = g.state_hi; rng[(size_t)n + i] = g.state_lo; rng[2 * (size_t)n + i] = g.inc_hi; rng[3 * (size_t)n + i] = g.inc_lo;
}

// ----------------------------------------------------------------------------- reset
template <int KIND, typename T>
__device__ __forceinline__ void reset_one(const Segment& seg, const uint8_t* mask, int i, unsigned int gseq) {
  typedef Traits<KIND> Tr;
  if (mask != nullptr && mask[i] == 0) {
    // not reset: with a fused gather attached the current row still has to reach the new slot
    if (seg.gth.n_peers > 0) {
      float o[Tr::D];
#pragma unroll
      for (int k = 0; k < Tr::D; ++k) o[k] = seg.obs[(size_t)i * Tr::D + k];
      gather_store_row<Tr::D>(seg.gth, gseq, (size_t)(seg.global_offset + i), o);
    }
    return;
  }
  T p[Tr::P];
  load_rows<KIND, T>(seg, i, Tr::P_STEP, Tr::P, p);
  Pcg64 g = load_rng(seg.rng, seg.n, i);
  pcg64_skip<Tr::GYM_DRAWS>(g);  // gymnasium's own (discarded) draws
  T s[Tr::S];
  float o[Tr::D];
  env_reset<KIND, T>(s, p, g, o);
  store_rng_state(seg.rng, seg.n, i, g);
  StateIO<T, Tr::S>::store(seg.state, i, s);
  store_obs<Tr::D>(seg.obs, (size_t)i, o);
  if (seg.gth.n_peers > 0) gather_store_row<Tr::D>(seg.gth, gseq, (size_t)(seg.global_offset + i), o);
  seg.elapsed[i] = 0;
  if (KIND == KIND_CARTPOLE) seg.sbt[i] = 0;
  seg.reward[i] = 0.0f;
  seg.terminated[i] = 0;
  seg.truncated[i] = 0;
}

template <int KIND, typename T>
__global__ void __launch_bounds__(kBlock) reset_kernel(const __grid_constant__ Segment seg, const uint8_t* mask) {
  const unsigned int gseq = gather_begin(seg.gth);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < seg.n) reset_one<KIND, T>(seg, mask, i, gseq);
  gather_epilogue_immediate(seg.gth, gseq);  // resets always push their own rows (no previous launch to defer to)
}

// ------------------------------------------------------------------------------ step
// Programmatic dependent launch (sm_90+): a step kernel lets the next launch in the stream start
// scheduling at once (launch_dependents) and, when it is itself launched with the programmatic
// serialization attribute, loads what the previous launch does not write (context rows, actions)
// before it waits for the previous grid to complete and flush (wait). Both instructions are no-ops
// for ordinary launches. This hides the ~1.5 us launch latency between back-to-back single-step
// launches, which is most of a 65 536-env step.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int KIND, typename T, bool CHK>
__device__ __forceinline__ void step_one(const Segment& seg, const void* actions, int i, const StepCheck& chk,
                                         unsigned int gseq) {
  typedef Traits<KIND> Tr;
  T p[Tr::P];
  load_rows<KIND, T>(seg, i, 0, Tr::P_STEP, p);
  const Action a = load_action(actions, seg.act_dtype, (long long)i);
  pdl_wait();  // everything below reads buffers the previous step launch writes
  T s[Tr::S];
  StateIO<T, Tr::S>::load(seg.state, i, s);
  int el = seg.elapsed[i];
  uint8_t sb = 0;
  if (KIND == KIND_CARTPOLE) sb = seg.sbt[i];
  uint64_t g0_hi = 0, g0_lo = 0;  // CHK: the PCG64 state before this step (when the step touches it)
  if (CHK) {
    StateIO<T, Tr::S>::store(chk.undo_state, i, s);
    chk.undo_elapsed[i] = el;
    if (KIND == KIND_CARTPOLE) chk.undo_sbt[i] = sb;
#pragma unroll
    for (int k = 0; k < Tr::D; ++k) chk.undo_obs[(size_t)i * Tr::D + k] = seg.obs[(size_t)i * Tr::D + k];
    chk.undo_reward[i] = seg.reward[i];
    chk.undo_flags[i] = seg.terminated[i];
    chk.undo_flags[(size_t)seg.n + i] = seg.truncated[i];
    if (Tr::DISCRETE && chk.n_actions > 0 && (unsigned)a.i >= (unsigned)chk.n_actions) {
      *reinterpret_cast<volatile int*>(chk.bad_action) = i + 1;  // any one offender is enough
      chk.undo_rng_flag[i] = 0;
      return;  // this env does not step; the host rolls the others back after the sync
    }
  }

  Pcg64 g;
  bool rng_live = false;
  T noise = (T)0;
  if (KIND == KIND_ACROBOT) {
    if (p[AC_NOISE] > (T)0) {  // env RNG draw, as AcrobotEnv.step does
      g = load_rng(seg.rng, seg.n, i);
      if (CHK) { g0_hi = g.state_hi; g0_lo = g.state_lo; }
      rng_live = true;
      noise = (T)pcg64_uniform(g, -(double)p[AC_NOISE], (double)p[AC_NOISE]);
    }
  }
  float o[Tr::D];
  const StepOut so = env_step<KIND, T>(s, p, a, noise, sb, o);
  el += 1;
  const bool tr = seg.max_steps > 0 && el >= seg.max_steps;
  if (seg.autoreset != CARLB_AUTORESET_NONE && (so.terminated || tr)) {
    if (seg.final_obs != nullptr) store_obs<Tr::D>(seg.final_obs, (size_t)i, o);
    load_rows<KIND, T>(seg, i, Tr::P_STEP, Tr::P, p);
    if (!rng_live) {
      g = load_rng(seg.rng, seg.n, i);
      if (CHK) { g0_hi = g.state_hi; g0_lo = g.state_lo; }
    }
    rng_live = true;
    pcg64_skip<Tr::GYM_DRAWS>(g);
    env_reset<KIND, T>(s, p, g, o);
    el = 0;
    sb = 0;
  }
  if (CHK) {
    chk.undo_rng_flag[i] = rng_live ? 1 : 0;
    if (rng_live) { chk.undo_rng[i] = g0_hi; chk.undo_rng[(size_t)seg.n + i] = g0_lo; }
  }
  if (rng_live) store_rng_state(seg.rng, seg.n, i, g);
  StateIO<T, Tr::S>::store(seg.state, i, s);
  store_obs<Tr::D>(seg.obs, (size_t)i, o);
  if (seg.gth.n_peers > 0 && seg.gth.mode == GATHER_IMMEDIATE)
    gather_store_row<Tr::D>(seg.gth, gseq, (size_t)(seg.global_offset + i), o);
  seg.reward[i] = so.reward;
  seg.terminated[i] = so.terminated ? 1 : 0;
  seg.truncated[i] = tr ? 1 : 0;
  seg.elapsed[i] = el;
  if (KIND == KIND_CARTPOLE) seg.sbt[i] = sb;
  if (seg.host_obs != nullptr) {  // zero-copy mirrors in mapped host memory
    store_obs<Tr::D>(seg.host_obs, (size_t)i, o);
    seg.host_reward[i] = so.reward;
    seg.host_terminated[i] = so.terminated ? 1 : 0;
    seg.host_truncated[i] = tr ? 1 : 0;
  }
}

// DEFERRED push of the fused gather (common.cuh): the CTA carries one extra warp. Every compute thread
// first stores the row the PREVIOUS launch left in seg.obs into every rank's gathered buffer (posted
// NVLink stores: the thread does not wait for them) and arrives on a named barrier; the extra warp
// waits on that barrier and does the part that would stall: the system-scope fence that drains the CTA's
// peer stores, the CTA count and -- in the last CTA -- the flag publication. The physics of this launch
// runs meanwhile. Returns false in the threads of the publisher warp (which are done).
template <int D>
__device__ __forceinline__ bool deferred_push_prologue(const Segment& seg, unsigned int gseq, int nct, int i) {
  if ((int)threadIdx.x >= nct) {
    named_bar_sync(kBarPushed, nct + 32);
    // The thread that publishes (last CTA to arrive) also waits for every rank's flag of this push -- mid-kernel,
    // with nothing else to do -- so the completion of the grid implies that the gathered tensor is complete here;
    // the compute warps never see a barrier, a counter or a fence.
    if ((int)threadIdx.x == nct && gather_publish(seg.gth, gseq, gridDim.x)) gather_wait_all(seg.gth, gseq);
    return false;
  }
  if (i < seg.n) {
    float o[D];
    const float* src = seg.obs + (size_t)i * D;
    if (D % 4 == 0) {
#pragma unroll
      for (int k = 0; k < D; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + k);
        o[k] = v.x; o[k + 1] = v.y; o[k + 2] = v.z; o[k + 3] = v.w;
      }
    } else if (D % 2 == 0) {
#pragma unroll
      for (int k = 0; k < D; k += 2) {
        const float2 v = *reinterpret_cast<const float2*>(src + k);
        o[k] = v.x; o[k + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) o[k] = src[k];
    }
    gather_store_row<D>(seg.gth, gseq, (size_t)(seg.global_offset + i), o);
  }
  named_bar_arrive(kBarPushed, nct + 32);
  return true;
}

// Ahead of the deferred push (whose row store waits an L2 round trip for the row load and the push counter): pull
// the lines this thread's env reads first -- state row, context rows, counters -- towards L1, so that the two waits
// overlap instead of adding up.
template <int KIND, typename T>
__device__ __forceinline__ void prefetch_ctx(const Segment& seg, int i) {
  const T* ctx = static_cast<const T*>(seg.ctx);
#pragma unroll
  for (int r = 0; r < Traits<KIND>::P; ++r) asm volatile("prefetch.global.L1 [%0];" ::"l"(ctx + (size_t)r * seg.n + i));
}
template <int KIND, typename T>
__device__ __forceinline__ void prefetch_env(const Segment& seg, int i) {
  typedef Traits<KIND> Tr;
  const T* ctx = static_cast<const T*>(seg.ctx);
  asm volatile("prefetch.global.L1 [%0];" ::"l"(static_cast<const T*>(seg.state) + (size_t)i * Tr::S));
#pragma unroll
  for (int r = 0; r < Tr::P; ++r) asm volatile("prefetch.global.L1 [%0];" ::"l"(ctx + (size_t)r * seg.n + i));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(seg.elapsed + i));
  if (KIND == KIND_CARTPOLE) asm volatile("prefetch.global.L1 [%0];" ::"l"(seg.sbt + i));
}

template <int KIND, typename T>
__global__ void __launch_bounds__(kBlock + 32) step_kernel(const __grid_constant__ Segment seg, const void* actions) {
  pdl_launch_dependents();
  const bool deferred = seg.gth.n_peers > 0 && seg.gth.mode == GATHER_DEFERRED;  // never together with PDL
  const int nct = deferred ? (int)blockDim.x - 32 : (int)blockDim.x;             // compute threads per CTA
  const unsigned int gseq = gather_begin(seg.gth);
  const int i = blockIdx.x * nct + threadIdx.x;
  if (deferred && !deferred_push_prologue<Traits<KIND>::D>(seg, gseq, nct, i)) return;
  if (i < seg.n) step_one<KIND, T, false>(seg, actions, i, StepCheck{}, gseq);
  else pdl_wait();
  if (!deferred) gather_epilogue_immediate(seg.gth, gseq);
}

// The host-buffer step with in-kernel action validation and an undo log (StepCheck).
template <int KIND, typename T>
__global__ void __launch_bounds__(kBlock) step_checked_kernel(const __grid_constant__ Segment seg, const void* actions,
                                                              const __grid_constant__ StepCheck chk) {
  const unsigned int gseq = gather_begin(seg.gth);
  const int i = chk.first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < chk.first + chk.count) step_one<KIND, T, true>(seg, actions, i, chk, gseq);
  gather_epilogue_immediate(seg.gth, gseq);
  if (chk.done_word != nullptr) {
    // completion word in mapped host memory: the host polls it instead of synchronising the stream. CTAs
    // arrive with a device-scope release; the last one fences at system scope (its fence is cumulative over
    // every CTA's posted PCIe writes) and stores the ticket.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int prev = atomicAdd(chk.part_counter, 1u);
      if (prev == gridDim.x - 1) {
        __threadfence();
        *chk.part_counter = 0;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(chk.done_word) = chk.done_ticket;
      }
    }
  }
}

// Roll a checked step back: state / elapsed / steps-beyond flag of every env, PCG64 state where it moved.
template <int KIND, typename T>
__global__ void __launch_bounds__(kBlock) step_undo_kernel(const __grid_constant__ Segment seg,
                                                           const __grid_constant__ StepCheck chk) {
  typedef Traits<KIND> Tr;
  const int i = chk.first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= chk.first + chk.count) return;
  T s[Tr::S];
  StateIO<T, Tr::S>::load(chk.undo_state, i, s);
  StateIO<T, Tr::S>::store(seg.state, i, s);
  seg.elapsed[i] = chk.undo_elapsed[i];
  if (KIND == KIND_CARTPOLE) seg.sbt[i] = chk.undo_sbt[i];
  if (chk.undo_rng_flag[i] != 0) {
    seg.rng[i] = chk.undo_rng[i];
    seg.rng[(size_t)seg.n + i] = chk.undo_rng[(size_t)seg.n + i];
  }
  // the returned quantities: device buffers and (zero-copy mirrors) the caller's host arrays
  float o[Tr::D];
#pragma unroll
  for (int k = 0; k < Tr::D; ++k) o[k] = chk.undo_obs[(size_t)i * Tr::D + k];
  const float rw = chk.undo_reward[i];
  const uint8_t te = chk.undo_flags[i], tr = chk.undo_flags[(size_t)seg.n + i];
  store_obs<Tr::D>(seg.obs, (size_t)i, o);
  seg.reward[i] = rw; seg.terminated[i] = te; seg.truncated[i] = tr;
  if (seg.host_obs != nullptr) {
    store_obs<Tr::D>(seg.host_obs, (size_t)i, o);
    seg.host_reward[i] = rw; seg.host_terminated[i] = te; seg.host_truncated[i] = tr;
  }
}

// ----------------------------------------------------------------------- mixed batch
struct MixedParams {
  int n_seg;
  int block_start[CARLB_MAX_MIXED + 1];
  int precision[CARLB_MAX_MIXED];
  const void* actions[CARLB_MAX_MIXED];
  Segment seg[CARLB_MAX_MIXED];
};

template <typename T>
__device__ __forceinline__ void mixed_dispatch(const Segment& seg, const void* actions, int i) {
  switch (seg.kind) {
    case KIND_CARTPOLE: step_one<KIND_CARTPOLE, T, false>(seg, actions, i, StepCheck{}, 0u); break;
    case KIND_PENDULUM: step_one<KIND_PENDULUM, T, false>(seg, actions, i, StepCheck{}, 0u); break;
    case KIND_ACROBOT: step_one<KIND_ACROBOT, T, false>(seg, actions, i, StepCheck{}, 0u); break;
    case KIND_MOUNTAINCAR: step_one<KIND_MOUNTAINCAR, T, false>(seg, actions, i, StepCheck{}, 0u); break;
    default: step_one<KIND_MOUNTAINCAR_CONT, T, false>(seg, actions, i, StepCheck{}, 0u); break;
  }
}

// One launch for several homogeneous shards: whole blocks belong to one shard, so the kind
// switch is block-uniform (no divergence).
__global__ void __launch_bounds__(kBlock) mixed_step_kernel(const __grid_constant__ MixedParams mp) {
  pdl_launch_dependents();
  int k = 0;
  while (k + 1 < mp.n_seg && (int)blockIdx.x >= mp.block_start[k + 1]) ++k;
  const Segment& seg = mp.seg[k];
  const int i = ((int)blockIdx.x - mp.block_start[k]) * blockDim.x + threadIdx.x;
  if (i >= seg.n) return;
  if (mp.precision[k] == CARLB_F64) mixed_dispatch<double>(seg, mp.actions[k], i);
  else mixed_dispatch<float>(seg, mp.actions[k], i);
}

// --------------------------------------------------------------------- fused rollout
// REC = true: the common fast path -- in-kernel policy, all four trajectory sinks present: no
// per-iteration null checks, one running element offset for all four streams. REC = false: generic
// (optional sinks, optional given actions).
// Loop-invariant kernel parameters live in the constant bank; left alone, the compiler re-reads them
// with a uniform load (LDCU) in EVERY iteration and stalls the first consumer on it (ncu source page,
// profiles/r01k: ~14 % of the rollout kernel's stall samples). An empty asm makes the value opaque,
// which pins it in a register for the whole loop.
// (ptxas rematerialises a plain parameter read, so the value is routed through a warp shuffle -- once,
// before the loop -- which it has to treat as an ordinary per-thread register.)
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) { return __shfl_sync(__activemask(), v, 0); }
__device__ __forceinline__ int pin_reg(int v) { return (int)pin_reg((uint32_t)v); }
__device__ __forceinline__ size_t pin_reg(size_t v) {
  return ((size_t)pin_reg((uint32_t)(v >> 32)) << 32) | (size_t)pin_reg((uint32_t)v);
}
template <typename P> __device__ __forceinline__ P* pin_reg(P* v) {
  P* q = reinterpret_cast<P*>(pin_reg(reinterpret_cast<size_t>(v)));
  __builtin_assume(__isGlobal(q));  // keep STG (not generic ST) for the trajectory stores
  return q;
}

// AR: every terminated / truncated env is reset in the same step AND no env enters with its
// "steps beyond terminated" flag set, so that flag is identically 0 (CartPole's reward is the constant
// 1) -- the compiler drops its bookkeeping from the loop.
template <int KIND, typename T, bool REC, bool AR>
__device__ __forceinline__ void rollout_body(const Segment& seg, int i, int n_steps, uint64_t policy_seed,
                                             uint32_t step_base, const void* actions, const carlb_traj_t& traj,
                                             int refill_threshold, uint64_t* sv_slot, unsigned int gseq, int pdl_lead,
                                             float* o_fin) {
  typedef Traits<KIND> Tr;
  const int n = seg.n;
  const int max_steps = pin_reg(seg.max_steps);
  // programmatic dependent launch: the step at which this thread lets the NEXT launch of the stream begin to
  // schedule its CTAs (they block in griddepcontrol.wait until this grid has completed); < 0: never
  const int pdl_step = pdl_lead > 0 ? pin_reg(n_steps > pdl_lead ? n_steps - pdl_lead : 0) : -1;
  const size_t row_stride = pin_reg((size_t)n);
  n_steps = pin_reg(n_steps);
  step_base = pin_reg(step_base);
  float* const tj_obs = pin_reg(traj.obs);
  void* const tj_act = pin_reg(traj.actions);
  float* const tj_rew = pin_reg(traj.reward);
  uint8_t* const tj_done = pin_reg(traj.done);
  T s[Tr::S];
  StateIO<T, Tr::S>::load(seg.state, i, s);
  T p[Tr::P];
  load_rows<KIND, T>(seg, i, 0, Tr::P, p);
  // Batched resets touch the env's PCG64 stream only inside the (rare) refill pass, so the stream
  // stays in HBM/L2 there instead of occupying eight loop-carried registers; the generic loop and
  // Acrobot (per-step noise draws) keep it in registers.
  constexpr bool batch_resets = AR && KIND != KIND_ACROBOT;
  Pcg64 g{};
  if (!batch_resets) g = load_rng(seg.rng, n, i);
  int el = seg.elapsed[i];
  uint8_t sb = 0;
  if (KIND == KIND_CARTPOLE && !AR) sb = seg.sbt[i];
  const uint64_t gid = (uint64_t)(seg.global_offset + i);
  PolicyStream ps = policy_stream(policy_seed, gid);
  float o[Tr::D];
  StepOut so;
  so.reward = 0.0f;
  so.terminated = false;
  bool tr = false;
  // running row pointers: row (t, i) of a [K][n] buffer advances by n elements per step
  float* t_obs = traj.obs != nullptr ? traj.obs + (size_t)i * Tr::D : nullptr;
  int32_t* t_act_i = static_cast<int32_t*>(traj.actions);
  float* t_act_f = static_cast<float*>(traj.actions);
  if (traj.actions != nullptr) { t_act_i += i; t_act_f += i; }
  float* t_rew = traj.reward != nullptr ? traj.reward + i : nullptr;
  uint8_t* t_done = traj.done != nullptr ? traj.done + i : nullptr;
  const size_t esz = seg.act_dtype == CARLB_ACT_I64 ? 8 : (seg.act_dtype == CARLB_ACT_U8 ? 1 : 4);
  const unsigned char* a_in = (!REC && actions != nullptr) ? static_cast<const unsigned char*>(actions) + (size_t)i * esz : nullptr;
  size_t off = (size_t)i;  // REC: element offset of row (t, i) in the [K][n] trajectory streams
  // Batched pre-generation of the next reset state. A reset costs ~300 integer instructions (two
  // 128-bit jump-ahead multiplies + S PCG64 draws); under a random policy ~5% of the envs of a warp
  // reset each step, i.e. ~80% of the warp-iterations would execute that divergent path for one or
  // two lanes. Instead every lane keeps its NEXT reset state ready in registers; consuming it is a
  // few moves, and the warp regenerates the missing ones together once `refill_threshold` (16) lanes need one, so
  // the expensive path runs ~5x less often at several times the lane utilisation. The per-env
  // PCG64 stream is consumed in exactly the same order as a step-by-step run (the draw merely
  // happens earlier); a pre-generated but unused reset is rolled back at kernel exit. Acrobot with
  // torque noise interleaves per-step draws on the same stream, so it keeps the in-place path.
  const int kRefill = pin_reg(refill_threshold);
  const unsigned lanes = __activemask();
  bool have_next = false;
  T ns[Tr::S];
  float no[Tr::D];
#pragma unroll
  for (int k = 0; k < Tr::S; ++k) ns[k] = (T)0;
#pragma unroll
  for (int k = 0; k < Tr::D; ++k) no[k] = 0.0f;
#pragma unroll 1
  for (int t = 0; t < n_steps; ++t) {
    if (t == pdl_step) pdl_launch_dependents();
    Action a;
    if (REC) a = policy_action<KIND>(ps, step_base + (uint32_t)t);
    else a = (a_in != nullptr) ? load_action(a_in, seg.act_dtype, 0) : policy_action<KIND>(ps, step_base + (uint32_t)t);
    T noise = (T)0;
    if (KIND == KIND_ACROBOT) {
      if (p[AC_NOISE] > (T)0) noise = (T)pcg64_uniform(g, -(double)p[AC_NOISE], (double)p[AC_NOISE]);
    }
    if (AR) sb = 0;
    so = env_step<KIND, T>(s, p, a, noise, sb, o);
    el += 1;
    tr = max_steps > 0 && el >= max_steps;
    const bool need_reset = (AR || seg.autoreset != CARLB_AUTORESET_NONE) && (so.terminated || tr);
    if (batch_resets) {
      // refill when enough lanes have used theirs up -- or when ANY lane must reset right now without
      // one in hand: the warp would execute the divergent in-place path for that single lane anyway,
      // so every lane that lacks a pre-generated state makes one in the same pass
      const unsigned lacking = __ballot_sync(lanes, !have_next);
      const unsigned urgent = __ballot_sync(lanes, need_reset && !have_next);
      if ((__popc(lacking) >= kRefill || urgent != 0u) && !have_next) {
        Pcg64 gr = load_rng(seg.rng, n, i);
        sv_slot[0] = gr.state_hi; sv_slot[kBlock] = gr.state_lo;  // shared memory: the roll-back point
        pcg64_skip<Tr::GYM_DRAWS>(gr);
        env_reset<KIND, T>(ns, p, gr, no);
        store_rng_state(seg.rng, n, i, gr);
        have_next = true;
      }
    }
    if (need_reset) {
      if (!batch_resets) {  // in-place reset: without batching (generic loop / Acrobot's interleaved noise draws)
        pcg64_skip<Tr::GYM_DRAWS>(g);
        env_reset<KIND, T>(s, p, g, o);
      } else {  // batching guarantees a pre-generated state here (the urgent refill above)
#pragma unroll
        for (int k = 0; k < Tr::S; ++k) s[k] = ns[k];
#pragma unroll
        for (int k = 0; k < Tr::D; ++k) o[k] = no[k];
        have_next = false;
      }
      el = 0;
      sb = 0;
    }
    if (REC) {
      store_obs<Tr::D>(tj_obs, off, o);
      if (Tr::DISCRETE) static_cast<int32_t*>(tj_act)[off] = a.i;
      else static_cast<float*>(tj_act)[off] = a.f;
      tj_rew[off] = so.reward;
      tj_done[off] = (uint8_t)((so.terminated ? 1 : 0) | (tr ? 2 : 0));
      off += row_stride;
    } else {
      if (t_obs != nullptr) { store_obs<Tr::D>(t_obs, 0, o); t_obs += (size_t)n * Tr::D; }
      if (traj.actions != nullptr) {
        if (Tr::DISCRETE) { *t_act_i = a.i; t_act_i += n; }
        else { *t_act_f = a.f; t_act_f += n; }
      }
      if (t_rew != nullptr) { *t_rew = so.reward; t_rew += n; }
      if (t_done != nullptr) { *t_done = (uint8_t)((so.terminated ? 1 : 0) | (tr ? 2 : 0)); t_done += n; }
      if (a_in != nullptr) a_in += (size_t)n * esz;
    }
  }
  if (batch_resets) {
    if (have_next) {  // roll back the reset that was pre-generated but never used
      seg.rng[i] = sv_slot[0]; seg.rng[(size_t)n + i] = sv_slot[kBlock];
    }
  } else {
    store_rng_state(seg.rng, n, i, g);
  }
  StateIO<T, Tr::S>::store(seg.state, i, s);
  if (n_steps > 0) {
    if (o_fin != nullptr) {  // deferred gather push: the kernel stores the row once the publisher warp has read the old one
#pragma unroll
      for (int k = 0; k < Tr::D; ++k) o_fin[k] = o[k];
    } else {
      store_obs<Tr::D>(seg.obs, (size_t)i, o);
    }
    if (seg.gth.n_peers > 0 && seg.gth.mode == GATHER_IMMEDIATE)
      gather_store_row<Tr::D>(seg.gth, gseq, (size_t)(seg.global_offset + i), o);
    seg.reward[i] = so.reward;
    seg.terminated[i] = so.terminated ? 1 : 0;
    seg.truncated[i] = tr ? 1 : 0;
  }
  seg.elapsed[i] = el;
  if (KIND == KIND_CARTPOLE) seg.sbt[i] = sb;
}

// The publisher warp of a rollout launch in DEFERRED gather mode (warp specialisation): it alone reads the rows the
// PREVIOUS launch left in seg.obs (this CTA's nct rows: 512 contiguous bytes per load instruction) and pushes them
// into every rank's gathered buffer, then releases the compute warps' final obs store (named barrier), fences, counts
// the CTA in and -- in the last CTA -- publishes the push and waits for the other ranks' flags. The compute warps run
// the physics meanwhile; all they ever see of the gather is one barrier before their last store, passed long before.
template <int D>
__device__ __forceinline__ void rollout_publisher(const Segment& seg, unsigned int gseq, int nct) {
  const int lane = (int)threadIdx.x - nct;
  const int base = blockIdx.x * nct;
  constexpr int kRows = kBlock / 32;
  float rows[kRows][D];
  bool valid[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int j = lane + 32 * r;
    valid[r] = j < nct && base + j < seg.n;
    if (valid[r]) {
      const float* src = seg.obs + (size_t)(base + j) * D;
      if (D % 4 == 0) {
#pragma unroll
        for (int k = 0; k < D; k += 4) {
          const float4 v = *reinterpret_cast<const float4*>(src + k);
          rows[r][k] = v.x; rows[r][k + 1] = v.y; rows[r][k + 2] = v.z; rows[r][k + 3] = v.w;
        }
      } else if (D % 2 == 0) {
#pragma unroll
        for (int k = 0; k < D; k += 2) {
          const float2 v = *reinterpret_cast<const float2*>(src + k);
          rows[r][k] = v.x; rows[r][k + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) rows[r][k] = src[k];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    if (valid[r]) gather_store_row<D>(seg.gth, gseq, (size_t)(seg.global_offset + base + lane + 32 * r), rows[r]);
  // the stores above consumed the loaded values, so the old rows have been read: the compute warps may overwrite them
  named_bar_arrive(kBarObsRead, nct + 32);
  __syncwarp();  // the other lanes' row stores happen-before lane 0's fence
  if (lane == 0 && gather_publish(seg.gth, gseq, gridDim.x)) gather_wait_all(seg.gth, gseq);
}

template <int KIND, typename T, bool REC>
__global__ void __launch_bounds__(kBlock + 32) rollout_kernel(const __grid_constant__ Segment seg, int n_steps,
                                                              uint64_t policy_seed, uint32_t step_base, const void* actions,
                                                              const carlb_traj_t traj, int refill_threshold, int pdl_lead) {
  __shared__ uint64_t sv_sh[2 * kBlock];  // per-thread PCG64 state saved before a pre-generated reset
  const bool deferred = seg.gth.n_peers > 0 && seg.gth.mode == GATHER_DEFERRED;
  const int nct = deferred ? (int)blockDim.x - 32 : (int)blockDim.x;  // compute threads per CTA
  const int i = blockIdx.x * nct + threadIdx.x;
  if (pdl_lead > 0) {
    // launched with the programmatic-serialization attribute: this CTA may have started while the previous launch
    // of the stream was still in its last steps. What does not depend on that launch -- the read-only context rows
    // -- is pulled towards L1 now; everything else waits for the previous grid to complete and flush.
    if (i < seg.n && (int)threadIdx.x < nct) prefetch_ctx<KIND, T>(seg, i);
    pdl_wait();
  }
  const unsigned int gseq = gather_begin(seg.gth);
  if (deferred && (int)threadIdx.x >= nct) {
    rollout_publisher<Traits<KIND>::D>(seg, gseq, nct);
    return;
  }
  float o_fin[Traits<KIND>::D];
  if (i < seg.n) {
    // the env's PCG64 words are first needed a few steps into the loop (batched reset pre-generation: every warp
    // refills once per launch); pulling them towards L1 now takes their L2 round trip off that pass
    if (seg.autoreset != CARLB_AUTORESET_NONE) {
#pragma unroll
      for (int r = 0; r < 4; ++r) asm volatile("prefetch.global.L1 [%0];" ::"l"(seg.rng + (size_t)r * seg.n + i));
    }
    bool clean = seg.autoreset != CARLB_AUTORESET_NONE;  // warp-uniform choice of the specialised loop
    if (KIND == KIND_CARTPOLE) clean = __all_sync(__activemask(), clean && seg.sbt[i] == 0);
    uint64_t* sv_slot = &sv_sh[threadIdx.x];
    float* of = deferred ? o_fin : nullptr;
    if (clean) rollout_body<KIND, T, REC, true>(seg, i, n_steps, policy_seed, step_base, actions, traj, refill_threshold, sv_slot, gseq, pdl_lead, of);
    else rollout_body<KIND, T, REC, false>(seg, i, n_steps, policy_seed, step_base, actions, traj, refill_threshold, sv_slot, gseq, pdl_lead, of);
  }
  // ONE call site reached by every (compute) thread of the CTA: the barriers below are aligned and must not be
  // executed from divergent code (ragged tail warps)
  if (deferred) {
    named_bar_sync(kBarObsRead, nct + 32);  // the publisher warp arrived once its loads of the OLD rows had completed
    if (i < seg.n && n_steps > 0) store_obs<Traits<KIND>::D>(seg.obs, (size_t)i, o_fin);
  } else {
    gather_epilogue_immediate(seg.gth, gseq);
  }
}

// --------------------------------------------------------------------------- launchers
#define CARLB_DISPATCH_KIND_T(KINDV, PREC, CALL)                                          \
  switch ((KINDV) * 2 + ((PREC) == CARLB_F64 ? 1 : 0)) {                                  \
    case KIND_CARTPOLE * 2: { constexpr int K_ = KIND_CARTPOLE; typedef float T_; CALL; } break;          \
    case KIND_CARTPOLE * 2 + 1: { constexpr int K_ = KIND_CARTPOLE; typedef double T_; CALL; } break;     \
    case KIND_PENDULUM * 2: { constexpr int K_ = KIND_PENDULUM; typedef float T_; CALL; } break;          \
    case KIND_PENDULUM * 2 + 1: { constexpr int K_ = KIND_PENDULUM; typedef double T_; CALL; } break;     \
    case KIND_ACROBOT * 2: { constexpr int K_ = KIND_ACROBOT; typedef float T_; CALL; } break;            \
    case KIND_ACROBOT * 2 + 1: { constexpr int K_ = KIND_ACROBOT; typedef double T_; CALL; } break;       \
    case KIND_MOUNTAINCAR * 2: { constexpr int K_ = KIND_MOUNTAINCAR; typedef float T_; CALL; } break;    \
    case KIND_MOUNTAINCAR * 2 + 1: { constexpr int K_ = KIND_MOUNTAINCAR; typedef double T_; CALL; } break; \
    case KIND_MOUNTAINCAR_CONT * 2: { constexpr int K_ = KIND_MOUNTAINCAR_CONT; typedef float T_; CALL; } break; \
    case KIND_MOUNTAINCAR_CONT * 2 + 1: { constexpr int K_ = KIND_MOUNTAINCAR_CONT; typedef double T_; CALL; } break; \
    default: set_error("unknown classic env kind %d", (int)(KINDV)); return CARLB_ERR_INVALID;           \
  }

static inline int grid_for(int n) { return (n + kBlock - 1) / kBlock; }

int classic_seed(const carlb_env* env, uint64_t seed, cudaStream_t st) {
  seed_kernel<<<grid_for(env->n), kBlock, 0, st>>>(env->bufs.rng, env->n, seed, env->global_offset);
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

// Fused cross-GPU gather: describe this launch's push (slot / flag value come from device memory).
static void attach_gather(const carlb_env* env, Segment& seg, int launch_kind) {
  if (env->gather == nullptr) return;
  gather_fill(env->gather, &seg.gth, launch_kind);
}
// block size of a step / rollout launch: one extra (publisher) warp per CTA when the push is deferred
static inline int block_with_publisher(const Segment& seg, int compute_threads) {
  return compute_threads + ((seg.gth.n_peers > 0 && seg.gth.mode == GATHER_DEFERRED) ? 32 : 0);
}

int classic_reset(const carlb_env* env, const uint8_t* mask, cudaStream_t st) {
  Segment seg = make_segment(env, CARLB_ACT_I32);
  attach_gather(env, seg, GL_RESET);
  CARLB_DISPATCH_KIND_T(env->kind, env->precision, (reset_kernel<K_, T_><<<grid_for(env->n), kBlock, 0, st>>>(seg, mask)));
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

// Programmatic dependent launch is OPT-IN (CARLB_PDL=1). Measured on B200 (profiles/README.md,
// r01d A/B): inside a CUDA-graph replay of back-to-back 65 536-env step launches the programmatic
// edges cost more than they hide (4.19 us vs 3.31 us per launch), so plain launches are the default.
static bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CARLB_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

template <int KIND, typename T>
static cudaError_t launch_step_pdl(const Segment& seg, const void* actions, int n, cudaStream_t st) {
  if (!pdl_enabled() || seg.gth.n_peers > 0) {  // the fused gather reads device-side counters the previous launch writes
    step_kernel<KIND, T><<<grid_for(n), block_with_publisher(seg, kBlock), 0, st>>>(seg, actions);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid_for(n));
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, step_kernel<KIND, T>, seg, actions);
}

int classic_step(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm) {
  Segment seg = make_segment(env, act_dtype);
  attach_gather(env, seg, GL_CLASSIC_STEP);
  if (hm != nullptr) {
    seg.host_obs = hm->obs; seg.host_reward = hm->reward; seg.host_terminated = hm->terminated;
    seg.host_truncated = hm->truncated;
  }
  cudaError_t le = cudaSuccess;
  CARLB_DISPATCH_KIND_T(env->kind, env->precision, (le = launch_step_pdl<K_, T_>(seg, actions, env->n, st)));
  g_launches++;
  CARLB_CUDA_CHECK(le);
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

int classic_step_checked(const carlb_env* env, const void* actions, int act_dtype, cudaStream_t st, const HostMirrors* hm,
                         const StepCheck& chk) {
  Segment seg = make_segment(env, act_dtype);
  attach_gather(env, seg, GL_IMMEDIATE_ONLY);
  if (hm != nullptr) {
    seg.host_obs = hm->obs; seg.host_reward = hm->reward; seg.host_terminated = hm->terminated;
    seg.host_truncated = hm->truncated;
  }
  CARLB_DISPATCH_KIND_T(env->kind, env->precision,
                        (step_checked_kernel<K_, T_><<<grid_for(chk.count), kBlock, 0, st>>>(seg, actions, chk)));
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

int classic_step_undo(const carlb_env* env, cudaStream_t st, const StepCheck& chk, const HostMirrors* hm) {
  Segment seg = make_segment(env, CARLB_ACT_I32);
  if (hm != nullptr) {
    seg.host_obs = hm->obs; seg.host_reward = hm->reward; seg.host_terminated = hm->terminated;
    seg.host_truncated = hm->truncated;
  }
  CARLB_DISPATCH_KIND_T(env->kind, env->precision, (step_undo_kernel<K_, T_><<<grid_for(chk.count), kBlock, 0, st>>>(seg, chk)));
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

// Compute threads per CTA of a rollout launch: the preferred size unless the grid would then need more than ONE
// wave of resident CTAs -- a rollout is a long per-thread loop, a second (partial) wave doubles the launch. The
// publisher warp of the deferred gather push makes the CTAs bigger, which can cost a resident CTA per SM (r02e:
// 96-thread CTAs at 94 registers -> 6 per SM = 888 slots < 1 024 CTAs, 11.4 -> 19.2 us); twice the compute threads per
// CTA then halve the grid.
template <int KIND, typename T, bool REC>
static void launch_rollout(const carlb_env* env, const Segment& seg, int preferred, int extra, int n_steps, uint64_t policy_seed,
                           uint32_t step_base, const void* actions, const carlb_traj_t& tj, int refill, cudaStream_t st) {
  static int n_sms[64] = {};
  static int occ[2][3] = {};  // [extra != 0][block 64 / 128 / 32]: resident CTAs per SM of this instantiation
  const int dev = env->device & 63;
  if (n_sms[dev] == 0) cudaDeviceGetAttribute(&n_sms[dev], cudaDevAttrMultiProcessorCount, env->device);
  int block = preferred;
  const int candidates[2] = {preferred, 128};
  for (int c = 0; c < 2; ++c) {
    block = candidates[c] > preferred ? candidates[c] : preferred;
    const int slot = block == 64 ? 0 : (block == 128 ? 1 : 2);
    int& o = occ[extra != 0][slot];
    if (o == 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, rollout_kernel<KIND, T, REC>, block + extra, 0) != cudaSuccess) {
      cudaGetLastError();
      o = 1;
    }
    if ((long long)(env->n + block - 1) / block <= (long long)o * n_sms[dev]) break;
  }
  const int grid = (env->n + block - 1) / block;
  // CARLB_ROLLOUT_PDL=L (opt-in, default 0 = plain launches): launch with the programmatic-serialization attribute and
  // let every thread release the next launch of the stream L steps before its last one, so that launch latency, CTA
  // scheduling and the context prefetch of launch k+1 could overlap the tail of launch k. Measured on B200 inside
  // CUDA-graph trains of 20-step launches (profiles/r02j_pdl_sweep.txt): L = 1 gains 1 % (11.51 -> 11.36 us), larger
  // leads lose (the early CTAs take resident slots from the running grid: L = 8 13.1 us) -- not worth a default.
  static const int pdl_lead = [] {
    const char* e = getenv("CARLB_ROLLOUT_PDL");
    const int v = e != nullptr ? atoi(e) : 0;
    return v < 0 ? 0 : v;
  }();
  if (pdl_lead == 0) {
    rollout_kernel<KIND, T, REC><<<grid, block + extra, 0, st>>>(seg, n_steps, policy_seed, step_base, actions, tj, refill, 0);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(block + extra));
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, rollout_kernel<KIND, T, REC>, seg, n_steps, policy_seed, step_base, actions, tj, refill, pdl_lead);
}

int classic_rollout(const carlb_env* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                    int act_dtype, const carlb_traj_t* traj, cudaStream_t st) {
  Segment seg = make_segment(env, act_dtype);
  attach_gather(env, seg, GL_CLASSIC_STEP);
  carlb_traj_t tj{};
  if (traj != nullptr) tj = *traj;
  // 64-thread blocks: at N = 65 536 that is 1024 blocks ~ 6.9 per SM (better tail balance over
  // 148 SMs than 128-thread blocks for a long-running per-thread loop)
  // (CARLB_ROLLOUT_BLOCK / CARLB_ROLLOUT_REFILL override the block size and the batched-reset refill
  // threshold for A/B measurements)
  static const int kRolloutBlock = [] {
    const char* e = getenv("CARLB_ROLLOUT_BLOCK");
    const int v = e != nullptr ? atoi(e) : 64;
    return (v == 32 || v == 64 || v == 128) ? v : 64;
  }();
  static const int kRefillThreshold = [] {
    const char* e = getenv("CARLB_ROLLOUT_REFILL");
    const int v = e != nullptr ? atoi(e) : 16;  // sweep on B200 (profiles/r01j_sweep.txt): 4 -> 0.332, 8 -> 0.286, 16 -> 0.278 ms
    return (v >= 1 && v <= 32) ? v : 16;
  }();
  const bool rec = actions == nullptr && tj.obs != nullptr && tj.actions != nullptr && tj.reward != nullptr && tj.done != nullptr;
  const int extra = block_with_publisher(seg, 0);  // 32 when the push is deferred to a publisher warp
  if (rec) {
    CARLB_DISPATCH_KIND_T(env->kind, env->precision,
                          (launch_rollout<K_, T_, true>(env, seg, kRolloutBlock, extra, n_steps, policy_seed, step_base, actions,
                                                        tj, kRefillThreshold, st)));
  } else {
    CARLB_DISPATCH_KIND_T(env->kind, env->precision,
                          (launch_rollout<K_, T_, false>(env, seg, kRolloutBlock, extra, n_steps, policy_seed, step_base, actions,
                                                         tj, kRefillThreshold, st)));
  }
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

int classic_mixed_step(carlb_env* const* envs, const void* const* actions, const int* act_dtypes, int n_handles,
                       cudaStream_t st) {
  MixedParams mp{};
  mp.n_seg = n_handles;
  int blocks = 0;
  for (int k = 0; k < n_handles; ++k) {
    mp.block_start[k] = blocks;
    mp.seg[k] = make_segment(envs[k], act_dtypes[k]);
    mp.precision[k] = envs[k]->precision;
    mp.actions[k] = actions[k];
    blocks += grid_for(envs[k]->n);
  }
  mp.block_start[n_handles] = blocks;
  mixed_step_kernel<<<blocks, kBlock, 0, st>>>(mp);
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  return CARLB_OK;
}

}  // namespace carlb
