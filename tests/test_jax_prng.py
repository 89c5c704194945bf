"""The JAX PRNG stream of the reference's Brax resets (carl/envs/brax/wrappers.py:41,54-59,80-81), restated:
oracle (numpy) against PUBLISHED known answers, and the product's kernel source (carl_b200/csrc/rng.h compiled
by g++ through tests/hostcheck) against the oracle. The CUDA build is checked in tests/test_brax_parity_gpu.py."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import jax_prng as jp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "jax_prng_known_answers.json")))


@pytest.fixture(scope="module")
def hc():
    from tests.hostcheck.build_hostcheck import build

    return ctypes.CDLL(build())


def test_oracle_reproduces_the_random123_known_answers():
    for v in GOLD["threefry2x32_20_kat"]:
        c, k, want = ([int(x, 16) for x in v[n]] for n in ("counter", "key", "out"))
        y0, y1 = jp.threefry2x32(k, [c[0]], [c[1]])
        assert [int(y0[0]), int(y1[0])] == want


def test_oracle_reproduces_the_outputs_jax_documents_for_prngkey0():
    g = GOLD["jax_docs_prngkey0"]
    key = jp.prng_key(0)
    assert key.tolist() == [0, 0]
    assert jp.split(key).tolist() == g["split"]
    assert float(jp.normal(key, 1)[0]) == pytest.approx(g["normal_1"], abs=5e-9)      # printed with 8 digits
    assert float(jp.uniform(key, 1)[0]) == pytest.approx(g["uniform_scalar"], abs=5e-9)
    sub = jp.split(key)[1]
    assert float(jp.normal(sub, 1)[0]) == pytest.approx(g["subkey_normal_1"], abs=5e-8)


def test_kernel_source_threefry_known_answers(hc):
    out = (ctypes.c_uint32 * 2)()
    for v in GOLD["threefry2x32_20_kat"]:
        c, k, want = ([int(x, 16) for x in v[n]] for n in ("counter", "key", "out"))
        hc.hc_threefry2x32(k[0], k[1], c[0], c[1], out)
        assert [out[0], out[1]] == want


@pytest.mark.parametrize("n", [1, 2, 3, 6, 9, 14, 15, 17, 18, 64])
def test_kernel_source_split_uniform_normal_match_the_oracle(hc, n):
    rng = np.random.default_rng(n)
    hc.hc_jax_uniform.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
    for _ in range(20):
        k = rng.integers(0, 2**32, size=2, dtype=np.uint64).astype(np.uint32)
        ks = np.zeros((n, 2), dtype=np.uint32)
        hc.hc_jax_split(int(k[0]), int(k[1]), n, ks.ctypes.data_as(ctypes.c_void_p))
        np.testing.assert_array_equal(ks, jp.split(k, n))
        for lo, hi in ((0.0, 1.0), (-0.1, 0.1), (-0.005, 0.005), (-0.01, 0.01)):
            u = np.zeros(n, dtype=np.float32)
            hc.hc_jax_uniform(int(k[0]), int(k[1]), n, lo, hi, u.ctypes.data_as(ctypes.c_void_p))
            np.testing.assert_array_equal(u, jp.uniform(k, n, lo, hi))        # bit-exact
        z = np.zeros(n, dtype=np.float32)
        hc.hc_jax_normal(int(k[0]), int(k[1]), n, z.ctypes.data_as(ctypes.c_void_p))
        np.testing.assert_allclose(z, jp.normal(k, n), rtol=2e-6, atol=1e-7)   # erf_inv: last-ulp libm differences


def test_kernel_source_env_reset_key_chain(hc):
    out = (ctypes.c_uint32 * 2)()
    for seed, n_resets, batch, idx in [(0, 0, 1, 0), (0, 1, 1, 0), (0, 3, 1, 0), (0, 0, 8, 5), (7, 2, 4096, 4095), (2**33 + 5, 1, 3, 2)]:
        hc.hc_jax_env_reset_key(ctypes.c_uint64(seed), n_resets, batch, idx, out)
        assert [out[0], out[1]] == jp.env_reset_key(seed, n_resets, batch, idx).tolist()
    # seed 0, first reset of the unbatched shell: key2 of split(PRNGKey(0)) -- the documented pair
    hc.hc_jax_env_reset_key(ctypes.c_uint64(0), 0, 1, 0, out)
    assert [out[0], out[1]] == GOLD["jax_docs_prngkey0"]["split"][1]
