"""TEST INFRASTRUCTURE — numpy restatement of the JAX PRNG calls the reference's Brax envs make at reset.

JAX is not installable here (profiles/r02a_pip_install_attempt.txt), so this restates the PUBLISHED algorithm:
Threefry-2x32 with 20 rounds (Salmon et al., Random123) as `jax/_src/prng.py` wires it up -- `PRNGKey`, the
original (non-"partitionable") `threefry_random_bits` / `split` layout -- and `jax.random.uniform` / `normal`
(`jax/_src/random.py`; `normal` = sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1)) with XLA's float32 erf_inv
polynomial, xla/client/lib/math.cc). Call sites in the reference: carl/envs/brax/wrappers.py:41 (`seed(0)`),
:54-59 / :121-128 (`key1, key2 = jax.random.split(key)`; `env.reset(key2)`), :69-72, :80-81; brax 0.12.1
`VmapWrapper.reset` (`split(rng, batch_size)`) and `envs/<body>.py: reset` (`rng, rng1, rng2 = split(rng, 3)`,
`uniform(rng1, (q_size,), minval, maxval)`, `normal(rng2, (qd_size,))`).

PINNED by known answers (tests/golden/jax_prng_known_answers.json, tools/make_jax_prng_known_answers.py): the
Random123 KAT vectors of threefry2x32_20 and the outputs JAX's own documentation prints for PRNGKey(0).
Not part of the product.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32


def _rotl(x, r):
    return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(key, x0, x1):
    """Threefry-2x32-20 of the counter pairs (x0[i], x1[i]) under `key` = (k0, k1)."""
    k0, k1 = U32(key[0]), U32(key[1])
    x0 = np.asarray(x0, dtype=U32).copy()
    x1 = np.asarray(x1, dtype=U32).copy()
    ks = [k0, k1, U32(k0 ^ k1 ^ U32(0x1BD11BDA))]
    rot = [[13, 15, 26, 6], [17, 29, 16, 24]]
    with np.errstate(over="ignore"):
        x0 = (x0 + ks[0]).astype(U32)
        x1 = (x1 + ks[1]).astype(U32)
        for i in range(5):
            for r in rot[i % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(i + 2) % 3] + U32(i + 1)).astype(U32)
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=U32)


def random_bits(key, n: int) -> np.ndarray:
    cnt = np.arange(n, dtype=U32)
    if n % 2:
        cnt = np.concatenate([cnt, np.zeros(1, U32)])
    h = len(cnt) // 2
    y0, y1 = threefry2x32(key, cnt[:h], cnt[h:])
    return np.concatenate([y0, y1])[:n]


def split(key, num: int = 2) -> np.ndarray:
    return random_bits(key, 2 * num).reshape(num, 2)


def uniform(key, n: int, minval=0.0, maxval=1.0) -> np.ndarray:
    bits = random_bits(key, n)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, (f * (hi - lo)).astype(np.float32) + lo).astype(np.float32)


_A = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
      0.246640727, 1.50140941]
_B = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
      1.00167406, 2.83297682]


def erf_inv_f32(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p(-(x * x))).astype(np.float32)
    lt = w < np.float32(5)
    w = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3)).astype(np.float32)
    p = np.where(lt, np.float32(_A[0]), np.float32(_B[0])).astype(np.float32)
    for i in range(1, 9):
        p = (np.where(lt, np.float32(_A[i]), np.float32(_B[i])) + p * w).astype(np.float32)
    return (p * x).astype(np.float32)


def normal(key, n: int) -> np.ndarray:
    lo = np.nextafter(np.float32(-1), np.float32(0))
    return (np.float32(np.sqrt(2)) * erf_inv_f32(uniform(key, n, lo, 1.0))).astype(np.float32)


def env_reset_key(seed: int, n_resets: int, batch: int, env_index: int) -> np.ndarray:
    """Key that `Env.reset` receives for env `env_index` of a batch at the `n_resets`-th reset (0-based) of the gym
    shell (wrappers.py:54-59: the shell keeps key1, hands key2 on; VmapWrapper splits key2 over the batch)."""
    key = prng_key(seed)
    for _ in range(n_resets):
        key = split(key)[0]
    key2 = split(key)[1]
    return split(key2, batch)[env_index] if batch > 1 else key2


def brax_reset_draws(seed, n_resets, batch, env_index, nq, nqd, q_noise, qd_noise, qd_uniform):
    """(q noise, qd) of one env: `rng, rng1, rng2 = split(rng, 3)`; uniform(rng1, (nq,), -s, s); qd = s * normal(rng2)
    or uniform(rng2, (nqd,), -s, s). Returns (dq[nq], qd[nqd], rng) -- `rng` is what Reacher's target draw splits."""
    rng, rng1, rng2 = split(env_reset_key(seed, n_resets, batch, env_index), 3)
    dq = uniform(rng1, nq, -q_noise, q_noise)
    qd = uniform(rng2, nqd, -qd_noise, qd_noise) if qd_uniform else (np.float32(qd_noise) * normal(rng2, nqd)).astype(np.float32)
    return dq, qd, rng
