// Peer-store microbenchmark (2 GPUs, one process): how fast can SMs of GPU 0 write into GPU 1's memory over NVLink with
// (a) 16-byte st.global per thread and (b) TMA bulk stores (cp.async.bulk.global.shared::cta) of 2 KB per CTA
// iteration -- as sustained streams (64 MB) and as the short bursts the fused gather produces (1 MB, 7 MB)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_store_bench nvlink_store_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void st_kernel(float4* dst, const float4* src, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void bulk_kernel(unsigned char* dst, const unsigned char* src, size_t bytes, int chunk) {
  extern __shared__ __align__(128) unsigned char sm[];
  const size_t n_chunks = bytes / chunk;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const float4* s = reinterpret_cast<const float4*>(src + c * chunk);
    float4* d = reinterpret_cast<float4*>(sm);
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) d[i] = s[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(sa), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  const size_t maxb = 64ull << 20;
  unsigned char *src, *dst_peer, *dst_local;
  CK(cudaMalloc(&src, maxb));
  CK(cudaMalloc(&dst_local, maxb));
  CK(cudaMemset(src, 1, maxb));
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&dst_peer, maxb));
  CK(cudaSetDevice(0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t sizes[] = {1ull << 20, 7ull << 20, 64ull << 20};
  for (int target = 0; target < 2; ++target) {
    unsigned char* dst = target == 0 ? dst_peer : dst_local;
    for (size_t bytes : sizes) {
      for (int variant = 0; variant < 4; ++variant) {
        const int reps = bytes >= (64ull << 20) ? 20 : 200;
        float best = 1e9f;
        for (int trial = 0; trial < 3; ++trial) {
          CK(cudaDeviceSynchronize());
          CK(cudaEventRecord(e0));
          for (int r = 0; r < reps; ++r) {
            if (variant == 0) st_kernel<<<148 * 8, 256>>>((float4*)dst, (const float4*)src, bytes / 16);
            else if (variant == 1) bulk_kernel<<<148 * 4, 128, 2048>>>(dst, src, bytes, 2048);
            else if (variant == 2) bulk_kernel<<<148 * 4, 256, 8192>>>(dst, src, bytes, 8192);
            else CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice));
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        const char* names[] = {"st.v4 per thread", "bulk 2KB", "bulk 8KB", "cudaMemcpyAsync"};
        printf("%s  %5.0f MB  %-18s  %8.2f us per transfer  %7.1f GB/s\n", target == 0 ? "peer " : "local", bytes / 1048576.0, names[variant],
               best / reps * 1e3, bytes * (double)reps / (best * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
