// Random streams of the batched-step engine.
//
// * PCG64 + SeedSequence: the reference's classic-control resets draw from
//   `np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))` (gymnasium
//   `utils.seeding.np_random`, reached from carl/envs/carl_env.py:271 and consumed at
//   carl/envs/gymnasium/classic_control/carl_cartpole.py:51-61 etc.). To make device-side
//   (auto)resets bit-identical to that stream the same integer algorithms run per env:
//   one 128-bit LCG state + increment per env instance.
// * Philox4x32-10: counter-based stream for the synthetic random policy of the fused
//   rollout kernels (keyed by seed, global env id, step) -- independent of sharding.
//
// Everything here is integer work and must be bit-exact; tests compare against numpy
// itself (PCG64 / SeedSequence) and the Random123 known-answer vectors (Philox).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CARLB_HD __host__ __device__ __forceinline__
#else
#define CARLB_HD static inline
#endif

namespace carlb {

// ----------------------------------------------------------------------------- PCG64
struct Pcg64 {
  uint64_t state_hi, state_lo, inc_hi, inc_lo;
};

// 128-bit multiplier of PCG64 (PCG_DEFAULT_MULTIPLIER_128)
#define CARLB_PCG_MULT_HI 0x2360ED051FC65DA4ULL
#define CARLB_PCG_MULT_LO 0x4385DF649FCCF645ULL

CARLB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// state = state * MULT + inc  (mod 2^128)
CARLB_HD void pcg64_advance1(Pcg64& g) {
  const uint64_t lo = g.state_lo * CARLB_PCG_MULT_LO;
  uint64_t hi = mulhi64(g.state_lo, CARLB_PCG_MULT_LO) + g.state_hi * CARLB_PCG_MULT_LO +
                g.state_lo * CARLB_PCG_MULT_HI;
  const uint64_t nlo = lo + g.inc_lo;
  hi += g.inc_hi + (nlo < lo ? 1ULL : 0ULL);
  g.state_lo = nlo;
  g.state_hi = hi;
}

// 128 x 128 -> low 128 bits
CARLB_HD void mul128(uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint64_t& r_hi, uint64_t& r_lo) {
  r_lo = a_lo * b_lo;
  r_hi = mulhi64(a_lo, b_lo) + a_hi * b_lo + a_lo * b_hi;
}

// Jump ahead by k in {1, 2, 4} draws at once: state <- A_k * state + C_k * inc (mod 2^128), with
// A_k = MULT^k and C_k = 1 + MULT + ... + MULT^(k-1). Used for gymnasium's own reset draws, which
// CARL's reset consumes and discards (two 128-bit multiplies instead of k).
template <int K>
CARLB_HD void pcg64_skip(Pcg64& g) {
  static_assert(K == 1 || K == 2 || K == 4, "supported skips");
  if (K == 1) {
    pcg64_advance1(g);
    return;
  }
  const uint64_t a_hi = K == 2 ? 0x17bce35bdf69743cULL : 0xf4dd417327db7a9bULL;
  const uint64_t a_lo = K == 2 ? 0x529ed9eb20e0ae99ULL : 0xd194dfbe42d45771ULL;
  const uint64_t c_hi = K == 2 ? 0x2360ed051fc65da4ULL : 0x610e11a14b07e063ULL;
  const uint64_t c_lo = K == 2 ? 0x4385df649fccf646ULL : 0x817fa187adefba1cULL;
  uint64_t s_hi, s_lo, i_hi, i_lo;
  mul128(g.state_hi, g.state_lo, a_hi, a_lo, s_hi, s_lo);
  mul128(g.inc_hi, g.inc_lo, c_hi, c_lo, i_hi, i_lo);
  const uint64_t lo = s_lo + i_lo;
  g.state_hi = s_hi + i_hi + (lo < s_lo ? 1ULL : 0ULL);
  g.state_lo = lo;
}

// numpy pcg64_next64: step, then XSL-RR output of the new state
CARLB_HD uint64_t pcg64_next64(Pcg64& g) {
  pcg64_advance1(g);
  const uint64_t x = g.state_hi ^ g.state_lo;
  const unsigned rot = (unsigned)(g.state_hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

// numpy next_double: 53 random bits
CARLB_HD double pcg64_next_double(Pcg64& g) {
  return (double)(pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0);
}

// Generator.uniform(low, high) == low + (high - low) * next_double  (numpy random_uniform)
CARLB_HD double pcg64_uniform(Pcg64& g, double low, double high) {
  return low + (high - low) * pcg64_next_double(g);
}

// pcg_setseq_128_srandom_r(initstate, initseq)
CARLB_HD void pcg64_srandom(Pcg64& g, uint64_t st_hi, uint64_t st_lo, uint64_t seq_hi, uint64_t seq_lo) {
  g.state_hi = 0;
  g.state_lo = 0;
  g.inc_hi = (seq_hi << 1) | (seq_lo >> 63);
  g.inc_lo = (seq_lo << 1) | 1ULL;
  pcg64_advance1(g);
  const uint64_t lo = g.state_lo + st_lo;
  g.state_hi = g.state_hi + st_hi + (lo < g.state_lo ? 1ULL : 0ULL);
  g.state_lo = lo;
  pcg64_advance1(g);
}

// ----------------------------------------------------------------------- SeedSequence
// numpy.random.SeedSequence(entropy=<non-negative int < 2^64>).generate_state(4, uint64)
// followed by PCG64's seeding (initstate = words 0,1 ; initseq = words 2,3).
CARLB_HD uint32_t ss_hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= 0x931e8875u;
  value *= hash_const;
  value ^= value >> 16;
  return value;
}
CARLB_HD uint32_t ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
  r ^= r >> 16;
  return r;
}

CARLB_HD void pcg64_seed_from_int(Pcg64& g, uint64_t entropy) {
  uint32_t ent[2] = {(uint32_t)(entropy & 0xffffffffu), (uint32_t)(entropy >> 32)};
  const int n_ent = ent[1] != 0u ? 2 : 1;
  uint32_t pool[4];
  uint32_t hc = 0x43b0d7e5u;
  for (int i = 0; i < 4; ++i) pool[i] = ss_hashmix(i < n_ent ? ent[i] : 0u, hc);
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
  // generate_state(8 x uint32)
  uint32_t out[8];
  uint32_t hb = 0x8b51f9ddu;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= 0x58f38dedu;
    v *= hb;
    v ^= v >> 16;
    out[i] = v;
  }
  const uint64_t w0 = (uint64_t)out[0] | ((uint64_t)out[1] << 32);
  const uint64_t w1 = (uint64_t)out[2] | ((uint64_t)out[3] << 32);
  const uint64_t w2 = (uint64_t)out[4] | ((uint64_t)out[5] << 32);
  const uint64_t w3 = (uint64_t)out[6] | ((uint64_t)out[7] << 32);
  // numpy pcg64_set_seed(seed={w0,w1}, inc={w2,w3}): high word first
  pcg64_srandom(g, w0, w1, w2, w3);
}

// ---------------------------------------------------------------------- Philox4x32-10
struct Philox4 {
  uint32_t v[4];
};

CARLB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

CARLB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// uniform in [0,1) with 24 random bits (exactly representable in fp32)
CARLB_HD float u32_to_unit_float(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Synthetic-policy draw for (seed, global env id, step): 4 words
CARLB_HD Philox4 policy_draw(uint64_t seed, uint64_t env_id, uint32_t step) {
  return philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), step, 0x43415242u /*"CARB"*/,
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}

// ------------------------------------------------------------- JAX threefry2x32 PRNG
// The Brax envs the reference steps draw their reset noise from JAX's default PRNG
// (carl/envs/brax/wrappers.py:41,54-59,69-72,80-81 -> brax `Env.reset(rng)`): Threefry-2x32 (20 rounds,
// Random123) behind `jax.random.PRNGKey / split / uniform / normal` (jax/_src/prng.py, jax/_src/random.py; the
// original, non-"partitionable" bit layout). Pinned by the Random123 known answers and by the outputs JAX's own
// documentation prints (tests/golden/jax_prng_known_answers.json).
struct JaxKey {
  uint32_t k0, k1;
};

CARLB_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

CARLB_HD void threefry2x32(JaxKey key, uint32_t x0, uint32_t x1, uint32_t* y0, uint32_t* y1) {
  const uint32_t ks[3] = {key.k0, key.k1, key.k0 ^ key.k1 ^ 0x1BD11BDAu};
  const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 5; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *y0 = x0;
  *y1 = x1;
}

// jax.random.PRNGKey(seed): (high word, low word)
CARLB_HD JaxKey jax_prng_key(uint64_t seed) { return JaxKey{(uint32_t)(seed >> 32), (uint32_t)seed}; }

// Element `idx` of `threefry_random_bits(key, 32, (n,))`: the counts 0..n-1 (padded with one 0 when n is odd) are
// cut into two halves that go through the block cipher pairwise; the outputs are concatenated.
CARLB_HD uint32_t jax_random_bits(JaxKey key, uint32_t n, uint32_t idx) {
  const uint32_t h = (n + 1u) / 2u;            // half length of the (padded) count array
  const uint32_t j = idx < h ? idx : idx - h;  // position inside its half
  const uint32_t c1 = (h + j < n) ? h + j : 0u;  // second half: counts h.. (the pad element is 0)
  uint32_t y0, y1;
  threefry2x32(key, j, c1, &y0, &y1);
  return idx < h ? y0 : y1;
}

// jax.random.split(key, num)[which]
CARLB_HD JaxKey jax_split(JaxKey key, uint32_t num, uint32_t which) {
  return JaxKey{jax_random_bits(key, 2u * num, 2u * which), jax_random_bits(key, 2u * num, 2u * which + 1u)};
}

// jax.random.uniform(key, (n,), float32, minval, maxval)[idx]  (bit-exact)
CARLB_HD float jax_uniform(JaxKey key, uint32_t n, uint32_t idx, float minval, float maxval) {
  const uint32_t bits = jax_random_bits(key, n, idx);
  const uint32_t fb = (bits >> 9) | 0x3F800000u;
  float f;
#if defined(__CUDA_ARCH__)
  f = __uint_as_float(fb);
  f = __fsub_rn(f, 1.0f);
  const float v = __fadd_rn(__fmul_rn(f, __fsub_rn(maxval, minval)), minval);  // XLA: multiply, then add (no contraction across ops)
#else
  memcpy(&f, &fb, sizeof(f));
  f = f - 1.0f;
  volatile float prod = f * (maxval - minval);
  const float v = prod + minval;
#endif
  return v > minval ? v : minval;
}

// XLA's float32 erf_inv (Giles' single-precision polynomial, xla/client/lib/math.cc ErfInv32)
CARLB_HD float xla_erf_inv_f32(float x) {
  float w = -log1pf(-x * x);
  const bool lt = w < 5.0f;
  w = lt ? w - 2.5f : sqrtf(w) - 3.0f;
  float p = lt ? 2.81022636e-08f : -0.000200214257f;
  p = (lt ? 3.43273939e-07f : 0.000100950558f) + p * w;
  p = (lt ? -3.5233877e-06f : 0.00134934322f) + p * w;
  p = (lt ? -4.39150654e-06f : -0.00367342844f) + p * w;
  p = (lt ? 0.00021858087f : 0.00573950773f) + p * w;
  p = (lt ? -0.00125372503f : -0.0076224613f) + p * w;
  p = (lt ? -0.00417768164f : 0.00943887047f) + p * w;
  p = (lt ? 0.246640727f : 1.00167406f) + p * w;
  p = (lt ? 1.50140941f : 2.83297682f) + p * w;
  return p * x;
}

// jax.random.normal(key, (n,), float32)[idx] = sqrt(2) * erf_inv(uniform(key, (n,), minval=nextafter(-1, 0), maxval=1)).
// The uniform draw is bit-exact; log1p / the polynomial may differ from XLA's in the last ulp.
CARLB_HD float jax_normal(JaxKey key, uint32_t n, uint32_t idx) {
  const float lo = -0.99999994f;  // nextafter(-1, 0) in float32
  return 1.41421356f * xla_erf_inv_f32(jax_uniform(key, n, idx, lo, 1.0f));
}

// The key `Env.reset` receives for env `env_index` of a batch of `batch` envs at the `n_resets`-th reset (0-based)
// of the gym shell: the shell keeps key1 of `key1, key2 = split(key)` and hands key2 on at every reset
// (wrappers.py:54-59,121-128); VmapWrapper.reset splits that over the batch (batch == 1: the unbatched shell).
CARLB_HD JaxKey jax_env_reset_key(uint64_t seed, uint32_t n_resets, uint32_t batch, uint32_t env_index) {
  JaxKey key = jax_prng_key(seed);
  for (uint32_t r = 0; r < n_resets; ++r) key = jax_split(key, 2u, 0u);
  const JaxKey key2 = jax_split(key, 2u, 1u);
  return batch > 1u ? jax_split(key2, batch, env_index) : key2;
}

}  // namespace carlb
