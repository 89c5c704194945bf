// Does Blackwell's packed FP32 (FMUL2 / FFMA2 through __fmul2_rn / __fadd2_rn, exact per lane) shrink the small-vector
// helpers of the Brax kernels? A rotate + cross chain, scalar vs packed (x, y) pairs:
//   for v in "" -DPACKED; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false $v -cubin -o pk.cubin \
//       tools/microbench/packed_fp32_probe.cu && cuobjdump -sass pk.cubin | grep -cE "FMUL|FADD|FFMA|MOV"; done
// Result (CUDA 12.9): scalar 100 FMUL + 76 FADD = 176; packed 44 FMUL + 36 FADD + 12 FMUL2 + 16 FFMA2 + 53 MOV = 161:
// the register-pair shuffles that cross / rotate need eat most of the saving (-9 %), see DESIGN.md (d).
#include <cuda_runtime.h>
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
#ifdef PACKED
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return v3(r.x, r.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y)); return v3(r.x, r.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(s, s)); return v3(r.x, r.y, a.z * s); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  // (a.y b.z - a.z b.y, a.z b.x - a.x b.z, a.x b.y - a.y b.x)
  float2 p = __fmul2_rn(make_float2(a.y, a.z), make_float2(b.z, b.x));
  float2 q = __fmul2_rn(make_float2(a.z, a.x), make_float2(b.y, b.z));
  float2 r = __fadd2_rn(p, make_float2(-q.x, -q.y));
  return v3(r.x, r.y, a.x * b.y - a.y * b.x);
}
#else
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
#endif
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Q4 { float w, x, y, z; };
__device__ __forceinline__ V3 rotate(V3 v, Q4 q) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = q.w;
  return (u * (2.0f * dot(u, v))) + (v * (s * s - dot(u, u))) + (cross(u, v) * (2.0f * s));
}
__global__ void k(const float* in, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 v = v3(in[i], in[i + n], in[i + 2 * n]);
  Q4 q; q.w = in[i + 3 * n]; q.x = in[i + 4 * n]; q.y = in[i + 5 * n]; q.z = in[i + 6 * n];
  V3 acc = v;
  for (int t = 0; t < 4; ++t) { acc = rotate(acc, q); acc = cross(acc, v) + acc; }
  out[i] = acc.x; out[i + n] = acc.y; out[i + 2 * n] = acc.z;
}
