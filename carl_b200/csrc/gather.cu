// Fused cross-GPU observation gather over NVLink (P2P stores + flag signalling).
//
// The path's only exchange step is the all-gather of the observation tensor. Instead
// of launching a collective after the step kernel, every rank owns a symmetric buffer
//     obs[2][n_global][D] | flags[MAX_PEERS] | block_counter
// allocated with cudaMalloc and mapped into the other ranks' processes through CUDA IPC. The step /
// reset / rollout kernels store each observation row straight into EVERY rank's buffer (slot =
// launch parity) while they compute, then the last CTA publishes "launch k done" into every peer's
// flag word (release, system scope). A consumer only needs `carlb_gather_wait`: a one-warp kernel
// that spins (acquire, system scope) until all ranks have published launch k -- no NCCL call, no
// host synchronisation, the transfer overlaps the physics.
//
// Double buffering makes the overwrite safe: rank B can only run launch k+2 (which reuses slot k)
// after it has seen every rank's flag k+1, and rank A publishes k+1 only after its own stream has
// finished consuming slot k.
#include <cuda_runtime.h>
#include <string.h>

#include <new>

#include "engine.h"

struct carlb_gather {
  int device = 0, rank = 0, world = 1, obs_dim = 0;
  long long n_global = 0;
  size_t slot_floats = 0;
  unsigned char* base[CARLB_MAX_PEERS] = {};  // base[r]: rank r's allocation mapped in this process
  bool opened[CARLB_MAX_PEERS] = {};
  unsigned int launches = 0;  // obs-producing launches issued so far
  carlb_env* attached[CARLB_MAX_MIXED] = {};  // handles whose kernels write into this gather
  int n_attached = 0;
};

namespace carlb {

// called by carlb_env_destroy: forget a handle that goes away before its gather
void gather_forget_env(carlb_gather* g, carlb_env* env) {
  for (int i = 0; i < g->n_attached; ++i)
    if (g->attached[i] == env) g->attached[i] = nullptr;
}

static size_t flags_offset(const carlb_gather* g) { return 2 * g->slot_floats * sizeof(float); }
static size_t total_bytes(const carlb_gather* g) {
  return flags_offset(g) + (CARLB_MAX_PEERS + 8) * sizeof(unsigned int);
}

__global__ void gather_wait_kernel(const unsigned int* flags, int world, unsigned int value) {
  const int r = threadIdx.x;
  if (r < world) {
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
    } while ((int)(v - value) < 0);
  }
}

// Called by the launchers: fills the peer fields of a kernel segment for the next obs-producing launch.
void gather_fill(carlb_gather* g, int* n_peers, float** peer_obs, unsigned int** peer_flags, unsigned int* signal_value,
                 unsigned int** block_counter) {
  const unsigned int k = g->launches++;
  const size_t slot = (k & 1u) * g->slot_floats;
  *n_peers = g->world;
  for (int r = 0; r < g->world; ++r) {
    peer_obs[r] = reinterpret_cast<float*>(g->base[r]) + slot;
    peer_flags[r] = reinterpret_cast<unsigned int*>(g->base[r] + flags_offset(g)) + g->rank;
  }
  *signal_value = k + 1;
  *block_counter = reinterpret_cast<unsigned int*>(g->base[g->rank] + flags_offset(g)) + CARLB_MAX_PEERS;
}

}  // namespace carlb

using namespace carlb;

extern "C" {

int carlb_gather_create(int device, int rank, int world, int64_t n_global, int obs_dim, carlb_gather_t** out) {
  if (out == nullptr || world < 1 || world > CARLB_MAX_PEERS || rank < 0 || rank >= world || n_global <= 0 || obs_dim <= 0) {
    set_error("carlb_gather_create: bad arguments (rank %d of %d, n_global %lld, obs_dim %d)", rank, world,
              (long long)n_global, obs_dim);
    return CARLB_ERR_INVALID;
  }
  carlb_gather* g = new (std::nothrow) carlb_gather();
  if (g == nullptr) return CARLB_ERR_STATE;
  g->device = device; g->rank = rank; g->world = world; g->obs_dim = obs_dim; g->n_global = n_global;
  g->slot_floats = ((size_t)n_global * obs_dim + 63) / 64 * 64;
  CARLB_CUDA_CHECK(cudaSetDevice(device));
  void* p = nullptr;
  CARLB_CUDA_CHECK(cudaMalloc(&p, total_bytes(g)));
  CARLB_CUDA_CHECK(cudaMemset(p, 0, total_bytes(g)));
  CARLB_CUDA_CHECK(cudaDeviceSynchronize());
  g->base[rank] = static_cast<unsigned char*>(p);
  *out = g;
  return CARLB_OK;
}

int carlb_gather_export(carlb_gather_t* g, void* handle64) {
  if (g == nullptr || handle64 == nullptr) {
    set_error("carlb_gather_export: null argument");
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
  if (g == nullptr || handle64 == nullptr || peer_rank < 0 || peer_rank >= g->world || peer_rank == g->rank) {
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
  if (env->gather != nullptr) {
    set_error("carlb_gather_attach: the handle already has a gather attached");
    return CARLB_ERR_STATE;
  }
  if (g->n_attached >= CARLB_MAX_MIXED) {
    set_error("carlb_gather_attach: too many handles attached");
    return CARLB_ERR_STATE;
  }
  g->attached[g->n_attached++] = env;
  env->gather = g;
  return CARLB_OK;
}

int carlb_gather_wait(carlb_gather_t* g, int lag, void* stream, float** gathered) {
  if (g == nullptr || gathered == nullptr || lag < 0 || lag > 1) {
    set_error("carlb_gather_wait: null argument or lag not in {0, 1}");
    return CARLB_ERR_INVALID;
  }
  if (g->launches <= (unsigned int)lag) {
    set_error("carlb_gather_wait: only %u observation-producing launches issued, lag %d", g->launches, lag);
    return CARLB_ERR_STATE;
  }
  CARLB_CUDA_CHECK(cudaSetDevice(g->device));
  const unsigned int k = g->launches - 1 - (unsigned int)lag;
  const unsigned int* flags = reinterpret_cast<const unsigned int*>(g->base[g->rank] + flags_offset(g));
  gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, g->world, k + 1);
  g_launches++;
  CARLB_CUDA_CHECK(cudaGetLastError());
  *gathered = reinterpret_cast<float*>(g->base[g->rank]) + (k & 1u) * g->slot_floats;
  return CARLB_OK;
}

int carlb_gather_destroy(carlb_gather_t* g) {
  if (g == nullptr) return CARLB_OK;
  for (int i = 0; i < g->n_attached; ++i)  // the handles must not keep pointing at freed buffers
    if (g->attached[i] != nullptr && g->attached[i]->gather == g) g->attached[i]->gather = nullptr;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();  // no kernel may still be storing into the buffers that are unmapped below
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank) continue;
    if (g->opened[r] && g->base[r]) cudaIpcCloseMemHandle(g->base[r]);
  }
  if (g->base[g->rank]) cudaFree(g->base[g->rank]);
  delete g;
  return CARLB_OK;
}

}  // extern "C"
