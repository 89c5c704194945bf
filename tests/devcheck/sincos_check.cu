// Device-side check (test infrastructure, built and run by tests/test_classic_parity_gpu.py on the GPU
// box): carlb::m_sincos's small-angle branch must be BIT-identical to CUDA's sincosf for EVERY float
// in (-0.78, 0.78), and identical outside it. Prints "mismatches <count> checked <count>".
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../carl_b200/csrc/physics_classic.h"

__global__ void check(uint32_t lo, uint32_t hi, unsigned long long* bad) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (uint64_t b = (uint64_t)lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= hi; b += stride) {
    for (int sign = 0; sign < 2; ++sign) {
      const float x = __uint_as_float((uint32_t)b | (sign ? 0x80000000u : 0u));
      float s0, c0, s1, c1;
      sincosf(x, &s0, &c0);
      carlb::m_sincos(x, &s1, &c1);
      if (__float_as_uint(s0) != __float_as_uint(s1) || __float_as_uint(c0) != __float_as_uint(c1)) ++local;
    }
  }
  if (local) atomicAdd(bad, local);
}

int main() {
  unsigned long long* bad;
  cudaMalloc(&bad, sizeof(*bad));
  cudaMemset(bad, 0, sizeof(*bad));
  // every non-negative float bit pattern up to 1.0f (covers the whole fast branch, incl. denormals,
  // and the hand-over to sincosf just above 0.78), both signs
  const uint32_t lo = 0u, hi = 0x3f800000u;
  check<<<148 * 8, 256>>>(lo, hi, bad);
  unsigned long long h = 0;
  cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) { printf("cuda error\n"); return 2; }
  printf("mismatches %llu checked %llu\n", h, 2ull * ((unsigned long long)hi - lo + 1));
  return h == 0 ? 0 : 1;
}
