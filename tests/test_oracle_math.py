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


def test_integer_range_checks_equal_their_float_definitions():
    """ddgi_fastmath.cuh: regular_component / regular_origin as unsigned compares of the bit patterns, and
    their three-component forms, against the float comparisons they replace - on every exponent boundary,
    zeros, denormals, Inf, NaN and 2^22 random bit patterns."""
    import ctypes as C
    import util
    rng = np.random.default_rng(7)
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 2.0 ** -70, np.nextafter(np.float32(2.0 ** -70), np.float32(0)), 2.0 ** -60,
                        np.nextafter(np.float32(2.0 ** -60), np.float32(0)), 2.0, np.nextafter(np.float32(2.0), np.float32(3)), 2.0 ** 20,
                        np.nextafter(np.float32(2.0 ** 20), np.float32(0)), np.inf, -np.inf, np.nan, 1.0, -1.0, 0.5, 3e38, -3e38], dtype=np.float32)
    bits = rng.integers(0, 2 ** 32, size=3 * (1 << 21), dtype=np.uint64).astype(np.uint32)
    x = np.concatenate([special, -special, bits.view(np.float32)])
    x = np.ascontiguousarray(x[: (x.size // 3) * 3])
    n = x.size
    comp, org = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    dir3, org3 = np.zeros(n // 3, dtype=np.uint8), np.zeros(n // 3, dtype=np.uint8)
    hs = util.hostsim()
    hs.sim_regular_checks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    hs.sim_regular_checks(x.ctypes.data, n, comp.ctypes.data, org.ctypes.data, dir3.ctypes.data, org3.ctypes.data)
    with np.errstate(invalid="ignore"):
        ax = np.abs(x)
        want_comp = (ax >= np.float32(8.6736174e-19)) & (ax <= np.float32(2.0))
        want_org = (ax == 0) | ((ax >= np.float32(8.4703295e-22)) & (ax < np.float32(1048576.0)))
    assert np.array_equal(comp.astype(bool), want_comp)
    assert np.array_equal(org.astype(bool), want_org)
    assert np.array_equal(dir3.astype(bool), want_comp.reshape(-1, 3).all(axis=1))
    assert np.array_equal(org3.astype(bool), want_org.reshape(-1, 3).all(axis=1))


def test_face_normal_shortcut_equals_the_literal_normal():
    """ddgi_trace.cuh: face_normal_axis (no normalize() when one component of p - cell centre is the clear
    winner) against face_normal_unit on hit-like points (on a face, 1e-4 inside), exact edges and corners
    (ties), near ties a few ulp apart, and degenerate inputs."""
    import ctypes as C
    import util
    rng = np.random.default_rng(11)
    n = 400000
    cell = rng.integers(-200, 200, size=(n, 3)).astype(np.float32)
    p = cell - rng.random((n, 3)).astype(np.float32)
    axis = rng.integers(0, 3, size=n)
    side = rng.integers(0, 2, size=n)
    idx = np.arange(n)
    p[idx, axis] = cell[idx, axis] - side.astype(np.float32) - np.where(side == 1, np.float32(-1e-4), np.float32(1e-4))   # just inside a face
    # ties and near ties: copy one coordinate's offset onto another, then nudge by 0..3 ulp
    t = slice(0, n // 4)
    off = p[t, 0] - (cell[t, 0] - np.float32(0.5))
    sign = np.where(rng.integers(0, 2, size=off.size) == 1, np.float32(1), np.float32(-1))
    p[t, 1] = (cell[t, 1] - np.float32(0.5)) + sign * off
    for k in range(3):
        sel = slice(k * (n // 16), (k + 1) * (n // 16))
        for _ in range(k + 1):
            p[sel, 1] = np.nextafter(p[sel, 1], np.float32(1e9))
    p[-6:] = [[np.nan, 0, 0], [np.inf, 1, 1], [0, 0, 0], [1e30, 1e30, 1e30], [0.5, 0.5, 0.5], [-0.5, -0.5, -0.5]]
    cell[-2:] = [[1, 1, 1], [0, 0, 0]]   # p exactly the cell centre: the zero vector
    p, cell = np.ascontiguousarray(p), np.ascontiguousarray(cell)
    fast, lit = np.zeros((n, 3), dtype=np.float32), np.zeros((n, 3), dtype=np.float32)
    hs = util.hostsim()
    hs.sim_face_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hs.sim_face_normals(p.ctypes.data, cell.ctypes.data, n, fast.ctypes.data, lit.ctypes.data)
    assert np.array_equal(fast.view(np.uint32), lit.view(np.uint32))


def test_byte_unpack_equals_the_division_for_every_byte():
    """ddgi_math.cuh: unorm8_to_float (imageLoad of an rgba8 channel without the IEEE division) == float32(b) / 255 for
    all 256 bytes."""
    import ctypes as C
    import util
    out = np.zeros(256, dtype=np.float32)
    hs = util.hostsim()
    hs.sim_unorm8.argtypes = [C.c_void_p]
    hs.sim_unorm8(out.ctypes.data)
    want = np.arange(256, dtype=np.float32) / np.float32(255.0)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
