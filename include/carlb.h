/*
 * libcarlb -- C ABI of the B200-native batched-step engine for CARL's contextual
 * classic-control and Brax-locomotion environments.
 *
 * The reference (automl/CARL) has no native boundary: its seam for this path is the gymnasium
 * Env protocol that `CARLEnv` wraps (carl/envs/carl_env.py:245-342) plus the attribute pokes of
 * its family adapters. Each entry point below names the reference interface it replaces.
 * Plain pointers and sizes only; no torch types. Unless marked HOST, every pointer is a DEVICE
 * pointer into caller-owned memory (the Python host layer passes `tensor.data_ptr()`); the
 * library never allocates or frees per step and never synchronises the device except in the
 * *_host entry points. Calls are stream-ordered on the `stream` argument (a cudaStream_t passed
 * as void*; NULL = legacy default stream). A handle is not thread-safe; one handle per shard.
 *
 * All functions return CARLB_OK (0) or a negative error code; carlb_last_error() returns the
 * thread-local message of the last failure.
 */
#ifndef CARLB_H_
#define CARLB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARLB_ABI_VERSION 1
#define CARLB_MAX_PEERS 8
#define CARLB_MAX_MIXED 8
#define CARLB_MAX_PARTS 8

/* error codes */
#define CARLB_OK 0
#define CARLB_ERR_INVALID (-1)  /* bad argument (the host layer raises ValueError / AssertionError) */
#define CARLB_ERR_STATE (-2)    /* call order: buffers not bound, not seeded, ... (RuntimeError) */
#define CARLB_ERR_CUDA (-3)     /* a CUDA runtime call failed (RuntimeError) */

/* env kinds: one per reference class on the hot path */
#define CARLB_CARTPOLE 0          /* carl/envs/gymnasium/classic_control/carl_cartpole.py:11 */
#define CARLB_PENDULUM 1          /* .../carl_pendulum.py:11 */
#define CARLB_ACROBOT 2           /* .../carl_acrobot.py:11 */
#define CARLB_MOUNTAINCAR 3       /* .../carl_mountaincar.py:11 */
#define CARLB_MOUNTAINCAR_CONT 4  /* .../carl_mountaincarcontinuous.py:11 */
#define CARLB_BRAX_ANT 16         /* carl/envs/brax/carl_ant.py:14 */
#define CARLB_BRAX_HALFCHEETAH 17 /* carl/envs/brax/carl_halfcheetah.py:14 */
#define CARLB_BRAX_HOPPER 18      /* carl/envs/brax/carl_hopper.py:14 */
#define CARLB_BRAX_WALKER2D 19    /* carl/envs/brax/carl_walker2d.py:14 (same kernels, new table) */
#define CARLB_BRAX_INVERTED_PENDULUM 20        /* carl/envs/brax/carl_inverted_pendulum.py:9 (slide joint + new table) */
#define CARLB_BRAX_INVERTED_DOUBLE_PENDULUM 21 /* carl/envs/brax/carl_inverted_double_pendulum.py:9 */
#define CARLB_BRAX_REACHER 22                  /* carl/envs/brax/carl_reacher.py:9 (two-slide target body) */
#define CARLB_BRAX_HUMANOID 23                 /* carl/envs/brax/carl_humanoid.py:14 (stacked 2- / 3-dof hinges, 244-entry obs) */
#define CARLB_BRAX_HUMANOIDSTANDUP 24          /* carl/envs/brax/carl_humanoidstandup.py:9 */
#define CARLB_BRAX_PUSHER 25                   /* carl/envs/brax/carl_pusher.py:11 (capsule-vs-ball contact pairs) */

/* state / context precision of a handle */
#define CARLB_F32 0 /* throughput mode: fp32 state and context in HBM */
#define CARLB_F64 1 /* reference-precision mode (classic control only; the reference is float64) */

/* action element types accepted by step / rollout */
#define CARLB_ACT_I32 0
#define CARLB_ACT_I64 1
#define CARLB_ACT_U8 2
#define CARLB_ACT_F32 3

/* auto-reset modes */
#define CARLB_AUTORESET_NONE 0 /* classic-control reference behaviour: the caller resets */
#define CARLB_AUTORESET_SAME_STEP 1
/* classic: done envs re-draw their state from their own PCG64 stream in the same launch
 * (obs = first obs of the new episode, terminal obs goes to final_obs), context unchanged.
 * Brax: brax AutoResetWrapper semantics -- state/obs replaced by the stored first state. */

typedef struct carlb_env carlb_env_t;
typedef struct carlb_gather carlb_gather_t;

typedef struct carlb_env_info {
  int kind;
  int state_words;       /* per-env state elements (T for classic, float for Brax) */
  int obs_dim;
  int act_dim;
  int act_discrete;      /* 1: Discrete(n_actions), 0: Box */
  int n_actions;
  int n_param_rows;      /* rows of the per-env kernel-parameter table ctx[P][n] */
  int n_step_rows;       /* rows read by step (the rest are reset-only) */
  int default_max_steps; /* gymnasium TimeLimit / brax EpisodeWrapper length */
  int gym_reset_draws;   /* gymnasium's own reset draws that CARL's reset discards */
  float act_low, act_high;
} carlb_env_info_t;

/* Caller-owned device buffers of one handle (n = n_envs, S = state_words, D = obs_dim,
 * P = n_param_rows, T = float or double by precision). */
typedef struct carlb_buffers {
  void* state;         /* T[n][S]   env state rows (classic) / float[n][S] link state (Brax) */
  void* ctx;           /* T[P][n]   per-env kernel parameters (context features, SoA) */
  int32_t* elapsed;    /* [n]       TimeLimit counter */
  uint8_t* sbt;        /* [n]       CartPole "steps_beyond_terminated is not None" flag */
  uint64_t* rng;       /* [4][n]    per-env PCG64: state_hi, state_lo, inc_hi, inc_lo */
  float* obs;          /* [n][D] */
  float* reward;       /* [n] */
  uint8_t* terminated; /* [n] */
  uint8_t* truncated;  /* [n] */
  float* final_obs;    /* [n][D] or NULL: terminal observation of auto-reset envs */
  void* first_state;   /* Brax: float[n][S] state stored at reset (AutoResetWrapper), else NULL */
  float* first_obs;    /* Brax: [n][D], else NULL */
  void* act_staging;   /* [n] x 8 bytes device scratch for carlb_env_step_host */
} carlb_buffers_t;

/* Optional trajectory sinks of a fused rollout (any pointer may be NULL). K = n_steps. */
typedef struct carlb_traj {
  float* obs;      /* [K][n][D] observation returned by step t */
  void* actions;   /* [K][n] int32 (discrete) or [K][n][A] float (continuous): action taken */
  float* reward;   /* [K][n] */
  uint8_t* done;   /* [K][n] bit0 = terminated, bit1 = truncated */
} carlb_traj_t;

int carlb_abi_version(void);
const char* carlb_last_error(void);

/* Static facts about an env kind (replaces reading gymnasium/brax spaces:
 * carl/envs/carl_env.py:77, carl/envs/brax/wrappers.py:45-52). */
int carlb_query_env(int kind, carlb_env_info_t* out);

/* Replaces `gymnasium.make(env_name)` (carl/envs/gymnasium/carl_gymnasium_env.py:63-64) and
 * `brax.envs.create(env_name, backend="spring", batch_size)` (carl/envs/brax/carl_brax_env.py:163-176)
 * for a shard of n_envs instances on `device`; global_offset = global id of local env 0. */
int carlb_env_create(int kind, int n_envs, int precision, int device, int64_t global_offset, carlb_env_t** out);
int carlb_env_destroy(carlb_env_t* env);

/* Hands the handle its buffers. Writing `ctx` is the batched `_update_context`
 * (carl_gymnasium_env.py:75-77 setattr loop; carl_brax_env.py:255-292 sys.replace). */
int carlb_env_bind(carlb_env_t* env, const carlb_buffers_t* bufs);

/* max_episode_steps: TimeLimit of gymnasium.make / EpisodeWrapper of brax.envs.create
 * (<= 0 keeps the kind's default); autoreset: CARLB_AUTORESET_*. */
int carlb_env_configure(carlb_env_t* env, int max_episode_steps, int autoreset);

/* `reset(seed=s)` seeding (carl/envs/carl_env.py:271 -> gymnasium np_random): env i gets
 * PCG64(SeedSequence(seed + global_offset + i)), bit-identical to numpy. */
int carlb_env_seed(carlb_env_t* env, uint64_t seed, void* stream);

/* `CARL<Env>.reset` for the envs with mask[i] != 0 (mask NULL = all): gymnasium's own reset draws
 * are consumed and discarded, then CARL's context-controlled draws produce the state
 * (carl_cartpole.py:44-66, carl_pendulum.py:41-65, carl_acrobot.py:71-115,
 * carl_mountaincar.py:53-85, carl_mountaincarcontinuous.py:50-82); Brax: `Ant.reset` etc. through
 * carl/envs/brax/wrappers.py:54-59 (q = init_q + noise, forward kinematics). Writes obs. */
int carlb_env_reset(carlb_env_t* env, const uint8_t* mask, void* stream);

/* Batched `CARLEnv.step(action)` (carl/envs/carl_env.py:321-342 -> gymnasium step /
 * carl/envs/brax/wrappers.py:62-67,74-78): one launch advances all n envs, writes
 * obs, reward, terminated, truncated (and final_obs / auto-reset). */
int carlb_env_step(carlb_env_t* env, const void* actions, int act_dtype, void* stream);

/* Same call with HOST buffers (what a gymnasium user holds: numpy in, numpy out): copies the
 * actions host->device, steps, copies obs/reward/terminated/truncated device->host and
 * synchronises `stream`. Host pointers should be page-locked for full PCIe rate. Any output
 * pointer may be NULL to skip that copy. */
int carlb_env_step_host(carlb_env_t* env, const void* actions_host, int act_dtype, float* obs_host,
                        float* reward_host, uint8_t* terminated_host, uint8_t* truncated_host, void* stream);

/* carlb_env_step_host plus the `assert self.action_space.contains(action)` of the gymnasium envs
 * (carl/envs/carl_env.py:339 -> gymnasium `step`): every discrete action must lie in [0, n_actions)
 * (n_actions <= 0: no check). With page-locked buffers the step kernel validates the actions itself while
 * it reads them (no host pass over the action array) and logs what it overwrites; if any action is
 * invalid the step is rolled back on the device -- the env is untouched, as in the reference -- and the
 * call returns CARLB_ERR_INVALID with the message "invalid action ...". Otherwise the check is a host
 * pass before the staged step. */
int carlb_env_step_host_checked(carlb_env_t* env, const void* actions_host, int act_dtype, int n_actions, float* obs_host,
                                float* reward_host, uint8_t* terminated_host, uint8_t* truncated_host, void* stream);

/* Split-batch stepping with HOST buffers (EnvPool-style send / recv; SB3's VecEnv step_async / step_wait):
 * the handle's envs are cut into n_parts contiguous parts, part p = envs [p*n/n_parts, (p+1)*n/n_parts).
 * _begin enqueues the step of ONE part on `stream` and returns at once. actions_host holds the PART's actions
 * (element 0 belongs to the part's first env); the result pointers are the FULL page-locked arrays (obs[n][D],
 * reward[n], ...), of which the kernel writes only the part's rows. The kernel reads the actions and writes the
 * results over PCIe itself and finally stores a completion word the host can poll;
 * _end waits for that word (no stream synchronisation) and, if an action of the part was invalid, rolls the
 * part back and returns CARLB_ERR_INVALID ("invalid action ..."). While the host consumes part p and prepares
 * its next actions, the other parts' results are crossing PCIe. Classic-control handles without a fused gather;
 * one step per part in flight; use a different stream per part. n_actions as in carlb_env_step_host_checked. */
int carlb_env_step_host_begin(carlb_env_t* env, int part, int n_parts, const void* actions_host, int act_dtype, int n_actions,
                              float* obs_host, float* reward_host, uint8_t* terminated_host, uint8_t* truncated_host,
                              void* stream);
int carlb_env_step_host_end(carlb_env_t* env, int part);

/* Host helper for the call above: copy a caller's (pageable) action array into the page-locked staging
 * block in one pass and, for discrete action dtypes with n_actions > 0, range-check it like the
 * `assert self.action_space.contains(action)` of the gymnasium envs the reference steps
 * (carl/envs/carl_env.py:339). CARLB_ERR_INVALID (message "invalid action ...") when a value is outside.
 * dst_pinned == NULL (or == src): range check only -- for action arrays that already are page-locked. */
int carlb_stage_actions(void* dst_pinned, const void* src, int64_t count, int act_dtype, int n_actions);

/* Fused K-step rollout (the `for t: env.step(policy(obs))` loop of a rollout worker in one
 * launch; state stays in registers). actions == NULL: synthetic random policy from
 * Philox4x32-10 keyed by (policy_seed, global env id, step_base + t); else actions[K][n]. */
int carlb_env_rollout(carlb_env_t* env, int n_steps, uint64_t policy_seed, uint32_t step_base, const void* actions,
                      int act_dtype, const carlb_traj_t* traj, void* stream);

/* Mixed-env batch: one launch steps several homogeneous shards (e.g. Pendulum + Acrobot). */
int carlb_mixed_step(carlb_env_t* const* envs, const void* const* actions, const int* act_dtypes, int n_handles,
                     void* stream);

/* Brax: upload the packed system table (HOST float[n_floats]) that the host layer builds from the
 * body model -- the batched stand-in for `mjcf.load(asset)` + `sys.replace(...)`
 * (carl/envs/brax/carl_brax_env.py:271-292). Layout: carl_b200/envs/brax_system.py /
 * carl_b200/csrc/physics_brax.h. stock_contact != 0 keeps the per-geom stock friction/elasticity
 * (context_mode="reference", where the reference's context never reaches the physics). */
int carlb_brax_set_system(carlb_env_t* env, const float* table, int n_floats, int stock_contact);

/* Brax arithmetic of the step / rollout kernels of a handle. CARLB_BRAX_STRICT (default): products and sums
 * rounded separately, reproduces the float32 restatement of the reference arithmetic to ~1e-6 per env-step (the
 * parity mode). CARLB_BRAX_FMA: a*b+c contracted into fused multiply-adds, as XLA does in the reference's own
 * compiled kernels -- fewer issued instructions, results at the float32 round-off floor of the algorithm against
 * a float64 evaluation (tests/test_brax_parity_gpu.py states the tolerance). */
#define CARLB_BRAX_STRICT 0
#define CARLB_BRAX_FMA 1
int carlb_brax_set_arithmetic(carlb_env_t* env, int arithmetic);

/* Brax reset-noise stream. CARLB_RESET_PHILOX (default): one Philox4x32-10 block per (seed, global env id, reset
 * count, index) -- invariant to sharding, not the reference's stream. CARLB_RESET_JAX: the reference's own stream --
 * JAX threefry2x32 keys exactly as carl/envs/brax/wrappers.py:41,54-59,69-72,80-81 and brax's VmapWrapper /
 * `Env.reset` consume them (PRNGKey(seed); per reset `key1, key2 = split(key)`; `split(key2, n_global)[env]`;
 * `rng, rng1, rng2 = split(rng, 3)`; uniform / normal draws): uniform draws bit-exact, normal draws to ~1e-6
 * (erf_inv). n_global = size of the whole batch (the reference's `batch_size`; 1 = the unbatched shell). */
#define CARLB_RESET_PHILOX 0
#define CARLB_RESET_JAX 1
int carlb_brax_set_reset_rng(carlb_env_t* env, int mode, int64_t n_global);

/* Brax parity-mode reset: `pipeline_init(q, qd)` (forward kinematics) from caller-supplied
 * generalized coordinates q[n][n_q], qd[n][n_qd] (DEVICE) instead of the noise draws of
 * `Ant.reset` etc. (the reference's JAX PRNG stream is not reproducible without JAX). */
int carlb_brax_reset_from_q(carlb_env_t* env, const uint8_t* mask, const float* q, const float* qd, void* stream);

/* Brax positional-goal epilogue: one `BraxWalkerGoalWrapper.step`
 * (carl/envs/brax/brax_walker_goal_wrapper.py:124-140) for every env instance, run on the handle's
 * current observation right after carlb_env_step, in float64 like the reference's NumPy:
 *   new = position + float32((obs[idx0], obs[idx1]) * dt);  reward = max(0, |goal - position| - |goal - new|);
 *   position = new;  reached = |goal - new| <= radius;  terminated |= reached;  success = reached.
 * All pointers are DEVICE memory: position[n][2] (in/out), goal[n][2], radius[n], reward[n] (out,
 * float64), success[n] (out, 0/1); `terminated` is the handle's bound flag buffer, OR-ed in place. */
int carlb_brax_goal_step(carlb_env_t* env, int idx0, int idx1, double dt, double* position, const double* goal,
                         const double* radius, double* reward, uint8_t* success, void* stream);

/* ---- fused cross-GPU observation gather (the path's one exchange step) ----------
 * One symmetric buffer per rank -- four slots of [n_global][obs_dim] floats, a flag word per rank and a few
 * control words -- mapped into every other rank's process (CUDA IPC, or symmetric memory the caller
 * provides). Once a gather is attached, every observation-producing launch of the handle (reset / step /
 * rollout) also stores ("pushes") its obs rows into every rank's buffer over NVLink and publishes its
 * push count; slot and flag value come from a counter in DEVICE memory, so the launches can be captured
 * into a CUDA graph and replayed. All ranks must issue the same sequence of observation-producing
 * launches. Replaces `VectorGymWrapper`'s batched obs return (carl/envs/brax/wrappers.py:136-146) for a
 * batch sharded over GPUs.
 *
 * CARLB_GATHER_SYNC (default): a launch pushes the rows it computes and its last CTA waits for every rank's
 *   push of the same launch -- when launch k has completed, the gathered tensor of observation k is complete.
 * CARLB_GATHER_PIPELINED: when launch k has completed, the gathered tensor of observation k-1 is complete;
 *   classic step / rollout launches push the previous launch's rows from a dedicated warp per CTA while
 *   they compute (the NVLink transfer, its fence and the flags overlap the physics), Brax launches push at
 *   their end and wait one push behind. */
#define CARLB_GATHER_SYNC 0
#define CARLB_GATHER_PIPELINED 1
int64_t carlb_gather_bytes(int64_t n_global, int obs_dim); /* size of one rank's symmetric block */
int carlb_gather_create(int device, int rank, int world, int64_t n_global, int obs_dim, carlb_gather_t** out);
/* Same, over memory the caller made symmetric (every block zeroed, carlb_gather_bytes() long, 16-byte
 * aligned): bases[r] = rank r's block as mapped in THIS process (HOST array of DEVICE pointers);
 * multicast_base = NVLS multicast alias of the blocks or NULL -- with it one `multimem.st` per row reaches
 * every rank instead of one store per rank. */
int carlb_gather_create_symmetric(int device, int rank, int world, int64_t n_global, int obs_dim, void* const* bases,
                                  void* multicast_base, carlb_gather_t** out);
int carlb_gather_export(carlb_gather_t* g, void* handle64 /* HOST, 64 bytes out */);
int carlb_gather_open(carlb_gather_t* g, int peer_rank, const void* handle64 /* HOST, 64 bytes */);
int carlb_gather_attach(carlb_gather_t* g, carlb_env_t* env); /* one handle per gather */
int carlb_gather_set_mode(carlb_gather_t* g, int mode);        /* CARLB_GATHER_*; between launches */
/* The gathered tensor of the handle's latest observation (lag = 0) or of the one before it (lag = 1), valid
 * for work enqueued on `stream` after this call. Whenever the mode already guarantees completeness nothing
 * is launched; PIPELINED with lag = 0 appends a flush launch (push of the current observation + wait). A
 * slot is overwritten by the fourth push after it. */
int carlb_gather_wait(carlb_gather_t* g, int lag, void* stream, float** gathered /* HOST out: DEVICE pointer */);
/* After replaying CUDA graphs that contain observation-producing launches: synchronises `stream` and
 * re-reads the device-side push counter so that carlb_gather_wait returns the right slot again. */
int carlb_gather_resync(carlb_gather_t* g, void* stream);
int carlb_gather_destroy(carlb_gather_t* g);

/* Counters of kernels launched through this library since load (bench `gpu_launches`). */
int64_t carlb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CARLB_H_ */
