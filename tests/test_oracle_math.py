"""The oracle's pinned math and RNG (oracle/ddgi_oracle.c PIN 8) and the engine's
restatement of the same functions (csrc/ddgi_math.cuh via tests/hostsim) agree bit
for bit, and both are within 1 fp32 ulp of libm."""
import ctypes as C

import numpy as np

import util
from oracle import oracle


def _ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def _args():
    rng = np.random.default_rng(7)
    x = np.concatenate([
        rng.uniform(-7, 7, 200000), rng.uniform(-1e4, 1e4, 200000), rng.uniform(0, 6.2831855, 200000),
        np.array([0.0, -0.0, 1e-30, 3.1415927, 6.2831855, 1e6, -1e6]),
    ]).astype(np.float32)
    return x


def test_pin_sincos_within_one_ulp_of_libm():
    lib = oracle.load()
    x = _args()
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    lib.orc_pin_sincos(x.ctypes.data, x.size, s.ctypes.data, c.ctypes.data)
    ref_s = np.sin(x.astype(np.float64)).astype(np.float32)
    ref_c = np.cos(x.astype(np.float64)).astype(np.float32)
    assert _ulp_diff(s, ref_s).max() <= 1
    assert _ulp_diff(c, ref_c).max() <= 1
    # and almost always exactly the correctly rounded value
    assert (s == ref_s).mean() > 0.999 and (c == ref_c).mean() > 0.999


def test_pin_acos_within_one_ulp_of_libm():
    lib = oracle.load()
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-1, 1, 400000), np.array([-1.0, 1.0, 0.0, 0.5, -0.5, 0.4999999, 0.9999999, -0.9999999, 1e-20])]).astype(np.float32)
    out = np.zeros_like(x)
    lib.orc_pin_acos(x.ctypes.data, x.size, out.ctypes.data)
    ref = np.arccos(x.astype(np.float64)).astype(np.float32)
    assert _ulp_diff(out, ref).max() <= 1
    assert (out == ref).mean() > 0.999
    bad = np.array([1.5, -2.0, np.nan], dtype=np.float32)
    o2 = np.zeros_like(bad)
    lib.orc_pin_acos(bad.ctypes.data, bad.size, o2.ctypes.data)
    assert np.isnan(o2).all()


def test_engine_math_is_bit_identical_to_the_oracle():
    lib = oracle.load()
    hs = util.hostsim()
    x = _args()
    s1, c1, s2, c2 = (np.zeros_like(x) for _ in range(4))
    lib.orc_pin_sincos(x.ctypes.data, x.size, s1.ctypes.data, c1.ctypes.data)
    hs.sim_pin_sincos(x.ctypes.data, x.size, s2.ctypes.data, c2.ctypes.data)
    assert np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    assert np.array_equal(c1.view(np.uint32), c2.view(np.uint32))
    y = np.random.default_rng(3).uniform(-1.001, 1.001, 300000).astype(np.float32)
    a1, a2 = np.zeros_like(y), np.zeros_like(y)
    lib.orc_pin_acos(y.ctypes.data, y.size, a1.ctypes.data)
    hs.sim_pin_acos(y.ctypes.data, y.size, a2.ctypes.data)
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32))


def test_wang_hash_and_xorshift_known_answers():
    """Hand-evaluated from assets/shaders/probe_pass.comp:45-71."""
    lib = oracle.load()

    def wang(seed):
        seed &= 0xFFFFFFFF
        seed = (seed ^ 61) ^ (seed >> 16)
        seed = (seed * 9) & 0xFFFFFFFF
        seed = seed ^ (seed >> 4)
        seed = (seed * 0x27D4EB2D) & 0xFFFFFFFF
        seed = seed ^ (seed >> 15)
        return seed

    for s in (0, 1, 2, 511, 65535, 8388607, 0xFFFFFFFF):
        assert lib.orc_wang_hash(s) == wang(s)
    out = np.zeros(64, dtype=np.float32)
    lib.orc_rand_sequence(12345, 64, out.ctypes.data)
    st = wang(12345)
    for i in range(64):
        st ^= (st << 13) & 0xFFFFFFFF
        st ^= st >> 17
        st ^= (st << 5) & 0xFFFFFFFF
        assert out[i] == np.float32(np.float32(st) / np.float32(4294967296.0))
    assert (out >= 0).all() and (out <= 1).all()


def test_hemisphere_direction_is_unit_and_above_the_surface():
    lib = oracle.load()
    for n in ((0, 1, 0), (1, 0, 0), (0, 0, -1), (0, -1, 0)):
        nn = np.array(n, dtype=np.float32)
        for seed in range(200):
            d = np.zeros(3, dtype=np.float32)
            lib.orc_hemisphere(nn.ctypes.data, seed, d.ctypes.data)
            assert abs(np.linalg.norm(d) - 1) < 1e-5
            assert np.dot(d, nn) >= -1e-6
