// Fused cross-GPU observation gather over NVLink (peer / multicast stores + flag signalling), host side.
//
// The path's only exchange step is the all-gather of the observation tensor. Instead of
// launching a collective after the step kernel, every rank owns a symmetric buffer
//     obs[4][n_global][D] | flags[MAX_PEERS] | ctrl[8]
// mapped into the other ranks' processes (CUDA IPC of a cudaMalloc block, or memory the caller made
// symmetric itself -- e.g. torch.distributed._symmetric_memory, which also yields an NVLS multicast
// alias so that ONE multimem.st reaches every rank). The step / reset / rollout kernels store their
// observation rows straight into EVERY rank's buffer ("push") and publish a per-rank push count; slot
// and flag value are read from DEVICE memory (ctrl[0]), so the launches are CUDA-graph capturable. The
// device side is in common.cuh (GatherDev).
//
// Two modes (carlb_gather_set_mode):
//   SYNC       every obs-producing launch pushes the rows it computes at its end and its last CTA waits
//              until every rank has published the same push: when launch k completes, the gathered tensor
//              of obs k is complete on this rank. The NVLink drain and the flag round trip are on the
//              critical path -- what a consumer needs that feeds obs k into step k + 1.
//   PIPELINED  when launch k completes, the gathered tensor of obs k-1 is complete. Classic step / rollout
//              launches carry a publisher warp per CTA that pushes the previous launch's rows while the
//              physics of launch k runs (transfer, fence and flags overlap the compute); Brax launches push
//              at their end and wait one push behind. carlb_gather_wait(lag = 0) appends a flush launch.
//
// Slot reuse: a push p reuses the slot of push p - 4. In the laxest case (PIPELINED, Brax) rank B's launch
// p starts after B's launch p-1 ended, which waited until every rank published push p-2, so rank A has
// started launch p-2 and therefore finished consuming push p-4 (its consumer of push j is stream-ordered
// before its launch j+2). Four slots cover every mode.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "engine.h"

namespace {
struct PushRec {
  long long push = -1;            // push index (slot = push % kGatherSlots)
  long long content = -1;         // version of the observation it carries
  long long complete_after = -1;  // complete on this rank once the launch with this index has completed
};
}  // namespace

struct carlb_gather {
  int device = 0, rank = 0, world = 1, obs_dim = 0, mode = CARLB_GATHER_SYNC;
  long long n_global = 0;
  size_t slot_floats = 0;
  unsigned char* base[CARLB_MAX_PEERS] = {};  // base[r]: rank r's allocation mapped in this process
  unsigned char* mc_base = nullptr;           // multicast alias of the allocation (or null)
  bool opened[CARLB_MAX_PEERS] = {};
  bool owns_memory = true;
  long long launches = 0;  // launches that pushed so far (== device-side ctrl[0] unless graphs were replayed)
  long long version = -1;  // version of the handle's current observation
  PushRec ring[carlb::kGatherSlots];
  carlb_env* attached = nullptr;  // the handle whose kernels write into this gather
};

namespace carlb {

// called by carlb_env_destroy: forget a handle that goes away before its gather
void gather_forget_env(carlb_gather* g, carlb_env* env) {
  if (g->attached == env) g->attached = nullptr;
}

static size_t slot_floats_for(long long n_global, int obs_dim) { return ((size_t)n_global * obs_dim + 63) / 64 * 64; }
static size_t flags_offset(const carlb_gather* g) { return (size_t)kGatherSlots * g->slot_floats * sizeof(float); }
static size_t total_bytes_for(long long n_global, int obs_dim) {
  return (size_t)kGatherSlots * slot_floats_for(n_global, obs_dim) * sizeof(float) + (CARLB_MAX_PEERS + 8) * sizeof(unsigned int);
}

static void fill_dev(const carlb_gather* g, GatherDev* d, int mode, int wait_lag) {
  *d = GatherDev{};
  d->n_peers = g->world;
  d->mode = mode;
  d->wait_lag = wait_lag;
  static const int debug = [] {
    const char* e = getenv("CARLB_GATHER_DEBUG");
    return e != nullptr ? atoi(e) : 0;
  }();
  d->debug = debug;
  d->slot_floats = g->slot_floats;
  for (int r = 0; r < g->world; ++r) {
    d->peer_base[r] = reinterpret_cast<float*>(g->base[r]);
    d->peer_flags[r] = reinterpret_cast<unsigned int*>(g->base[r] + flags_offset(g)) + g->rank;
  }
  if (g->mc_base != nullptr) {
    d->mc_base = reinterpret_cast<float*>(g->mc_base);
    d->mc_flag = reinterpret_cast<unsigned int*>(g->mc_base + flags_offset(g)) + g->rank;
  }
  unsigned int* local = reinterpret_cast<unsigned int*>(g->base[g->rank] + flags_offset(g));
  d->my_flags = local;
  d->ctrl = local + CARLB_MAX_PEERS;
}

// Called by the launchers: describes the push of the next obs-producing launch and records, on the host,
// which observation version it carries and when it is complete.
void gather_fill(carlb_gather* g, GatherDev* out, int launch_kind) {
  const long long p = g->launches++;
  const bool pipelined = g->mode == CARLB_GATHER_PIPELINED;
  PushRec rec;
  rec.push = p;
  if (pipelined && launch_kind == GL_CLASSIC_STEP && g->version >= 0) {
    // the publisher warps push what the previous launch left in the obs buffer; complete when this launch ends
    fill_dev(g, out, GATHER_DEFERRED, 0);
    rec.content = g->version;
    rec.complete_after = p;
    g->version += 1;
  } else if (pipelined && launch_kind == GL_BRAX) {
    // pushes its own rows at its end, waits for every rank's PREVIOUS push
    fill_dev(g, out, GATHER_IMMEDIATE, 1);
    g->version += 1;
    rec.content = g->version;
    rec.complete_after = p + 1;
    for (PushRec& r : g->ring)  // the previous push is complete once this launch has completed
      if (r.push == p - 1 && r.complete_after > p) r.complete_after = p;
  } else {
    fill_dev(g, out, GATHER_IMMEDIATE, 0);
    g->version += 1;
    rec.content = g->version;
    rec.complete_after = p;
  }
  g->ring[p % kGatherSlots] = rec;
}

// The flush launch (carlb_gather_wait with lag 0 in PIPELINED mode): an immediate push of the handle's
// current observation buffer, elementwise, with the in-kernel wait for every rank.
__global__ void __launch_bounds__(256) gather_flush_kernel(const __grid_constant__ GatherDev g, const float* obs, long long n_floats,
                                                           long long global_first) {
  const unsigned int seq = gather_begin(g);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_floats) gather_store_elem(g, seq, (size_t)(global_first + i), obs[i]);
  gather_epilogue_immediate(g, seq);
}

__global__ void gather_wait_kernel(const unsigned int* flags, int world, unsigned int value) {
  const int r = threadIdx.x;
  if (r < world)
    while ((int)(ld_acquire_sys_u32(flags + r) - value) < 0) {}
}

}  // namespace carlb

using namespace carlb;

static int gather_init(carlb_gather* g, int device, int rank, int world, int64_t n_global, int obs_dim) {
  g->device = device; g->rank = rank; g->world = world; g->obs_dim = obs_dim; g->n_global = n_global;
  g->slot_floats = slot_floats_for(n_global, obs_dim);
  return CARLB_OK;
}

extern "C" {

int64_t carlb_gather_bytes(int64_t n_global, int obs_dim) {
  if (n_global <= 0 || obs_dim <= 0) return 0;
  return (int64_t)total_bytes_for(n_global, obs_dim);
}

int carlb_gather_create(int device, int rank, int world, int64_t n_global, int obs_dim, carlb_gather_t** out) {
  if (out == nullptr || world < 1 || world > CARLB_MAX_PEERS || rank < 0 || rank >= world || n_global <= 0 || obs_dim <= 0) {
    set_error("carlb_gather_create: bad arguments (rank %d of %d, n_global %lld, obs_dim %d)", rank, world,
              (long long)n_global, obs_dim);
    return CARLB_ERR_INVALID;
  }
  carlb_gather* g = new (std::nothrow) carlb_gather();
  if (g == nullptr) return CARLB_ERR_STATE;
  gather_init(g, device, rank, world, n_global, obs_dim);
  CARLB_CUDA_CHECK(cudaSetDevice(device));
  void* p = nullptr;
  const size_t bytes = total_bytes_for(n_global, obs_dim);
  CARLB_CUDA_CHECK(cudaMalloc(&p, bytes));
  CARLB_CUDA_CHECK(cudaMemset(p, 0, bytes));
  CARLB_CUDA_CHECK(cudaDeviceSynchronize());
  g->base[rank] = static_cast<unsigned char*>(p);
  g->owns_memory = true;
  *out = g;
  return CARLB_OK;
}

int carlb_gather_create_symmetric(int device, int rank, int world, int64_t n_global, int obs_dim, void* const* bases,
                                  void* multicast_base, carlb_gather_t** out) {
  if (out == nullptr || bases == nullptr || world < 1 || world > CARLB_MAX_PEERS || rank < 0 || rank >= world ||
      n_global <= 0 || obs_dim <= 0) {
    set_error("carlb_gather_create_symmetric: bad arguments (rank %d of %d, n_global %lld, obs_dim %d)", rank, world,
              (long long)n_global, obs_dim);
    return CARLB_ERR_INVALID;
  }
  for (int r = 0; r < world; ++r)
    if (bases[r] == nullptr || ((uintptr_t)bases[r] & 15u)) {
      set_error("carlb_gather_create_symmetric: rank %d's base pointer is null or not 16-byte aligned", r);
      return CARLB_ERR_INVALID;
    }
  carlb_gather* g = new (std::nothrow) carlb_gather();
  if (g == nullptr) return CARLB_ERR_STATE;
  gather_init(g, device, rank, world, n_global, obs_dim);
  for (int r = 0; r < world; ++r) g->base[r] = static_cast<unsigned char*>(bases[r]);
  g->mc_base = static_cast<unsigned char*>(multicast_base);
  g->owns_memory = false;
  CARLB_CUDA_CHECK(cudaSetDevice(device));
  // the caller zeroes its block before the rendezvous; flags and control words must start at 0
  *out = g;
  return CARLB_OK;
}

int carlb_gather_export(carlb_gather_t* g, void* handle64) {
  if (g == nullptr || handle64 == nullptr || !g->owns_memory) {
    set_error("carlb_gather_export: null argument or caller-owned symmetric memory");
    return CARLB_ERR_INVALID;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CARLB_CUDA_CHECK(cudaSetDevice(g->device));
  CARLB_CUDA_CHECK(cudaIpcGetMemHandle(&h, g->base[g->rank]));
  memcpy(handle64, &h, sizeof(h));
  return CARLB_OK;
}

int carlb_gather_open(carlb_gather_t* g, int peer_rank, const void* handle64) {
  if (g == nullptr || handle64 == nullptr || peer_rank < 0 || peer_rank >= g->world || peer_rank == g->rank || !g->owns_memory) {
    set_error("carlb_gather_open: bad peer rank %d", peer_rank);
    return CARLB_ERR_INVALID;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  CARLB_CUDA_CHECK(cudaSetDevice(g->device));
  CARLB_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  g->base[peer_rank] = static_cast<unsigned char*>(p);
  g->opened[peer_rank] = true;
  return CARLB_OK;
}

int carlb_gather_attach(carlb_gather_t* g, carlb_env_t* env) {
  if (g == nullptr || env == nullptr) {
    set_error("carlb_gather_attach: null argument");
    return CARLB_ERR_INVALID;
  }
  for (int r = 0; r < g->world; ++r)
    if (g->base[r] == nullptr) {
      set_error("carlb_gather_attach: rank %d's buffer has not been opened", r);
      return CARLB_ERR_STATE;
    }
  carlb_env_info_t info;
  carlb_query_env(env->kind, &info);
  if (info.obs_dim != g->obs_dim || env->device != g->device) {
    set_error("carlb_gather_attach: obs_dim / device mismatch");
    return CARLB_ERR_INVALID;
  }
  if (env->gather == g) return CARLB_OK;
  if (env->gather != nullptr || g->attached != nullptr) {
    set_error("carlb_gather_attach: one handle per gather (every push carries the rows of ONE shard)");
    return CARLB_ERR_STATE;
  }
  if (env->global_offset < 0 || env->global_offset + env->n > g->n_global) {
    set_error("carlb_gather_attach: the handle's env range [%lld, %lld) lies outside the gathered tensor (%lld rows)",
              env->global_offset, env->global_offset + env->n, g->n_global);
    return CARLB_ERR_INVALID;
  }
  g->attached = env;
  env->gather = g;
  return CARLB_OK;
}

int carlb_gather_set_mode(carlb_gather_t* g, int mode) {
  if (g == nullptr || (mode != CARLB_GATHER_SYNC && mode != CARLB_GATHER_PIPELINED)) {
    set_error("carlb_gather_set_mode: null gather or unknown mode %d", mode);
    return CARLB_ERR_INVALID;
  }
  g->mode = mode;
  return CARLB_OK;
}

int carlb_gather_wait(carlb_gather_t* g, int lag, void* stream, float** gathered) {
  if (g == nullptr || gathered == nullptr || lag < 0 || lag > 1) {
    set_error("carlb_gather_wait: null argument or lag not in {0, 1}");
    return CARLB_ERR_INVALID;
  }
  if (g->version < (long long)lag) {
    set_error("carlb_gather_wait: only %lld observation-producing launches issued, lag %d", g->version + 1, lag);
    return CARLB_ERR_STATE;
  }
  CARLB_CUDA_CHECK(cudaSetDevice(g->device));
  const long long want = g->version - lag;
  const long long last_launch = g->launches - 1;
  const PushRec* best = nullptr;
  for (const PushRec& r : g->ring)
    if (r.content == want && r.push >= 0 && (best == nullptr || r.push > best->push)) best = &r;
  if (best != nullptr && best->complete_after > last_launch) {
    // pushed, but its completion is only implied by a later launch (PIPELINED Brax, latest observation):
    // a one-warp kernel waits for every rank's flag of that push
    const unsigned int* flags = reinterpret_cast<const unsigned int*>(g->base[g->rank] + flags_offset(g));
    gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, g->world, (unsigned int)(best->push + 1));
    g_launches++;
    CARLB_CUDA_CHECK(cudaGetLastError());
  }
  if (best == nullptr) {
    if (lag != 0 || g->attached == nullptr) {
      set_error("carlb_gather_wait: observation %lld is no longer (or not yet) in a slot", want);
      return CARLB_ERR_STATE;
    }
    // PIPELINED, latest observation not pushed yet: flush launch (immediate push of the obs buffer + wait)
    carlb_env* env = g->attached;
    const long long p = g->launches++;
    GatherDev d;
    fill_dev(g, &d, GATHER_IMMEDIATE, 0);
    const long long n_floats = (long long)env->n * g->obs_dim;
    const int grid = (int)((n_floats + 255) / 256);
    gather_flush_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, env->bufs.obs, n_floats, env->global_offset * g->obs_dim);
    g_launches++;
    CARLB_CUDA_CHECK(cudaGetLastError());
    PushRec rec;
    rec.push = p; rec.content = want; rec.complete_after = p;
    g->ring[p % kGatherSlots] = rec;
    best = &g->ring[p % kGatherSlots];
  }
  *gathered = reinterpret_cast<float*>(g->base[g->rank]) + (size_t)(best->push % kGatherSlots) * g->slot_floats;
  return CARLB_OK;
}

int carlb_gather_resync(carlb_gather_t* g, void* stream) {
  if (g == nullptr) {
    set_error("carlb_gather_resync: null gather");
    return CARLB_ERR_INVALID;
  }
  CARLB_CUDA_CHECK(cudaSetDevice(g->device));
  CARLB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  unsigned int seq = 0;
  const unsigned int* ctrl = reinterpret_cast<const unsigned int*>(g->base[g->rank] + flags_offset(g)) + CARLB_MAX_PEERS;
  CARLB_CUDA_CHECK(cudaMemcpy(&seq, ctrl + GCTRL_SEQ, sizeof(seq), cudaMemcpyDeviceToHost));
  const long long delta = (long long)seq - (long long)(unsigned int)g->launches;
  if (delta < 0) {
    set_error("carlb_gather_resync: the device has published %u pushes, fewer than the %lld launches issued", seq, g->launches);
    return CARLB_ERR_STATE;
  }
  if (delta == 0) return CARLB_OK;
  // graph replays pushed `delta` more times than the host saw: the replayed launches repeat the captured
  // pattern, so the last pushes carry the same relative observation versions, shifted by delta
  PushRec shifted[kGatherSlots];
  for (const PushRec& r : g->ring) {
    if (r.push < 0) continue;
    PushRec s = r;
    s.push += delta; s.content += delta; s.complete_after += delta;
    shifted[s.push % kGatherSlots] = s;
  }
  for (int i = 0; i < kGatherSlots; ++i) g->ring[i] = shifted[i];
  g->launches += delta;
  g->version += delta;
  return CARLB_OK;
}

int carlb_gather_destroy(carlb_gather_t* g) {
  if (g == nullptr) return CARLB_OK;
  if (g->attached != nullptr && g->attached->gather == g) g->attached->gather = nullptr;  // no dangling buffers
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();  // no kernel may still be storing into the buffers that are unmapped below
  if (g->owns_memory) {
    for (int r = 0; r < g->world; ++r) {
      if (r == g->rank) continue;
      if (g->opened[r] && g->base[r]) cudaIpcCloseMemHandle(g->base[r]);
    }
    if (g->base[g->rank]) cudaFree(g->base[g->rank]);
  }
  delete g;
  return CARLB_OK;
}

}  // extern "C"
