"""Pins the counter-based random stream: Philox4x32-10 of the oracle against (1) the Random123 known-answer vectors,
(2) NVIDIA's cuRAND host generator CURAND_RNG_PSEUDO_PHILOX4_32_10, (3) an independent pure-Python restatement of the
published algorithm; and the u64 -> f64 conversion of rand 0.8's Uniform::new_inclusive(0.0, 1.0)."""
import ctypes as C
import ctypes.util
import os

import numpy as np
import pytest

from oracle import oracle_py


def philox_py(ctr, key, rounds=10):
    m0, m1, w0, w1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(rounds):
        p0, p1 = m0 * c[0], m1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + w0) & 0xFFFFFFFF, (k[1] + w1) & 0xFFFFFFFF]
    return c


def oracle_philox(ctr, key):
    L = oracle_py.lib()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    L.oracle_philox4x32_10(c, k, o)
    return list(o)


# Random123 kat_vectors, philox4x32 10 rounds
KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,expect", KAT)
def test_random123_known_answers(ctr, key, expect):
    assert oracle_philox(ctr, key) == expect
    assert philox_py(ctr, key) == expect


def test_against_python_restatement_random_inputs():
    rng = np.random.default_rng(1)
    for _ in range(200):
        ctr = [int(x) for x in rng.integers(0, 2**32, 4)]
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        assert oracle_philox(ctr, key) == philox_py(ctr, key)


def _curand():
    for name in ("libcurand.so.10", "/usr/local/cuda/lib64/libcurand.so.10", "/usr/local/cuda/lib64/libcurand.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    return None


def test_against_curand_host_generator():
    """cuRAND's host Philox generator returns philox4x32_10(counter = 0, key = seed) as its first four outputs."""
    lib = _curand()
    if lib is None:
        pytest.skip("libcurand not found")
    CURAND_RNG_PSEUDO_PHILOX4_32_10 = 161
    for seed in (0, 1234, 0xDEADBEEFCAFEF00D, 2**63 + 5):
        gen = C.c_void_p()
        assert lib.curandCreateGeneratorHost(C.byref(gen), CURAND_RNG_PSEUDO_PHILOX4_32_10) == 0
        assert lib.curandSetPseudoRandomGeneratorSeed(gen, C.c_ulonglong(seed)) == 0
        out = (C.c_uint32 * 4)()
        assert lib.curandGenerate(gen, out, C.c_size_t(4)) == 0
        lib.curandDestroyGenerator(gen)
        assert list(out) == oracle_philox([0, 0, 0, 0], [seed & 0xFFFFFFFF, seed >> 32])


def test_uniform_conversion_matches_rand_0_8_new_inclusive():
    L = oracle_py.lib()
    scale = 1.0 + 2.0 ** -52
    # UniformFloat::new_inclusive(0, 1): max_rand = 1 - 2^-52, scale = 1 / max_rand, lowered until scale * max_rand <= 1
    max_rand = 1.0 - 2.0 ** -52
    assert 1.0 / max_rand == scale and scale * max_rand + 0.0 <= 1.0
    for x in (0, 1, 2**12, 2**64 - 1, 0x123456789ABCDEF0, 2**63):
        m = x >> 12
        assert L.oracle_uniform_from_u64(x) == (m * 2.0 ** -52) * scale
    assert L.oracle_uniform_from_u64(2**64 - 1) == 1.0  # inclusive upper end
    assert L.oracle_uniform_from_u64(0) == 0.0
