"""Integer streams must be bit-exact: the product's PCG64 / SeedSequence against numpy itself
(the reference's reset RNG, carl_env.py:271 -> gymnasium np_random) and Philox4x32-10 against the
Random123 known-answer vectors. The functions under test are the product's __host__ __device__
sources (carl_b200/csrc/rng.h) compiled by g++ (tests/hostcheck); the GPU suite re-checks the
device build through the C ABI."""
import ctypes

import numpy as np
import pytest

from tests.util import HostCheck


@pytest.fixture(scope="module")
def hc():
    return HostCheck()


@pytest.mark.parametrize("seed", [0, 1, 42, 65535, 123456789, 2**32 - 1, 2**32, 2**40 + 17, 2**63 + 5, 2**64 - 1])
def test_seedsequence_pcg64_matches_numpy(hc, seed):
    out = (ctypes.c_uint64 * 4)()
    hc.lib.hc_pcg64_seed(ctypes.c_uint64(seed), out)
    bg = np.random.PCG64(np.random.SeedSequence(seed))
    st = bg.state["state"]
    m = 2**64 - 1
    assert list(out) == [st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m]
    d = np.zeros(64)
    hc.lib.hc_pcg64_doubles(out, 64, d.ctypes.data_as(ctypes.c_void_p))
    np.testing.assert_array_equal(d, np.random.Generator(bg).random(64))


def test_uniform_matches_generator_uniform(hc):
    out = (ctypes.c_uint64 * 4)()
    hc.lib.hc_pcg64_seed(ctypes.c_uint64(7), out)
    d = np.zeros(8)
    hc.lib.hc_pcg64_doubles(out, 8, d.ctypes.data_as(ctypes.c_void_p))
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(7)))
    lo, hi = -0.1, 0.3
    np.testing.assert_array_equal(lo + (hi - lo) * d[:4], g.uniform(lo, hi, 4))
    np.testing.assert_array_equal(0.0 + (2.5 - 0.0) * d[4:5], [g.uniform(high=2.5)])


KAT = [  # Random123 kat_vectors: philox4x32-10
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(hc, ctr, key, want):
    C = (ctypes.c_uint32 * 4)(*ctr)
    K = (ctypes.c_uint32 * 2)(*key)
    O = (ctypes.c_uint32 * 4)()
    hc.lib.hc_philox(C, K, O)
    assert list(O) == want


def test_policy_actions_in_range_and_uniform(hc):
    ai, af = ctypes.c_int(), ctypes.c_float()
    counts = np.zeros(3, int)
    vals = []
    for i in range(3000):
        hc.lib.hc_policy_action(2, ctypes.c_uint64(5), ctypes.c_uint64(i), 3, ctypes.byref(ai), ctypes.byref(af))
        counts[ai.value] += 1
        hc.lib.hc_policy_action(1, ctypes.c_uint64(5), ctypes.c_uint64(i), 3, ctypes.byref(ai), ctypes.byref(af))
        vals.append(af.value)
    assert counts.min() > 850
    vals = np.asarray(vals)
    assert vals.min() >= -2.0 and vals.max() < 2.0 and abs(vals.mean()) < 0.1


def test_binary_policy_shift_register_equals_fresh_stream(hc):
    """CartPole's random policy spends one Philox bit per step and walks it with a shift register;
    whatever the order of the steps asked for (consecutive runs, unaligned launch boundaries, jumps
    backwards), the action of (seed, env, step) must be the bit a fresh stream returns."""
    rng = np.random.default_rng(0)
    runs = [np.arange(0, 300), np.arange(13, 13 + 200), np.arange(4294967290 - 40, 4294967290),
            rng.integers(0, 5000, size=400), np.concatenate([np.arange(95, 140), np.arange(20, 70), np.arange(127, 131)])]
    ai, af = ctypes.c_int(), ctypes.c_float()
    for env_id, steps in enumerate(runs):
        steps = np.ascontiguousarray(steps, dtype=np.uint32)
        out = np.zeros(len(steps), dtype=np.int32)
        hc.lib.hc_policy_sequence(ctypes.c_uint64(77), ctypes.c_uint64(env_id), steps.ctypes.data_as(ctypes.c_void_p),
                                  len(steps), out.ctypes.data_as(ctypes.c_void_p))
        for s_, got in zip(steps, out):
            hc.lib.hc_policy_action(0, ctypes.c_uint64(77), ctypes.c_uint64(env_id), ctypes.c_uint32(int(s_)), ctypes.byref(ai),
                                    ctypes.byref(af))
            assert ai.value == got, (env_id, int(s_))
        assert 0.3 < out.mean() < 0.7
