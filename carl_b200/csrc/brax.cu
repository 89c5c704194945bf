// placeholder until the spring pipeline lands (replaced below in this round)
#include "engine.h"
namespace carlb {
int brax_query(int, carlb_env_info_t*) { set_error("Brax kernels not built yet"); return CARLB_ERR_INVALID; }
int brax_create(carlb_env*) { return CARLB_ERR_INVALID; }
void brax_destroy(carlb_env*) {}
int brax_seed(const carlb_env*, uint64_t, cudaStream_t) { return CARLB_ERR_INVALID; }
int brax_reset(const carlb_env*, const uint8_t*, cudaStream_t) { return CARLB_ERR_INVALID; }
int brax_step(const carlb_env*, const void*, int, cudaStream_t) { return CARLB_ERR_INVALID; }
int brax_rollout(const carlb_env*, int, uint64_t, uint32_t, const void*, int, const carlb_traj_t*, cudaStream_t) { return CARLB_ERR_INVALID; }
int brax_set_tunables(carlb_env*, const float*, int) { return CARLB_ERR_INVALID; }
int brax_get_tunables(int, float*, int, int*) { return CARLB_ERR_INVALID; }
}
