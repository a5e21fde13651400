"""CPU tier: the engine's per-ray headers (csrc/ddgi_*.cuh) compiled for the host by
tests/hostsim — both the reference-order tracer and the wavefront state machine with its
exact-division / directed-rounding shortcuts — against the oracle, bit for bit.
This is the no-GPU regression gate for kernel work; the GPU tier repeats it through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import oracle

CFG = util.configs.CONFIGS


def _sim_probe_update(sc, rays, variant, k0=0, k1=None):
    hs = util.hostsim()
    W, H = sc.tex_size
    k1 = sc.num_rays if k1 is None else k1
    alb = np.zeros((H, W), dtype=np.uint32)
    f32 = np.zeros((H, W, 4), dtype=np.float32)
    lk = np.zeros(sc.num_rays, dtype=np.uint32)
    hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, k0, k1, variant, alb.ctypes.data, f32.ctypes.data, lk.ctypes.data, None)
    return alb, f32, lk


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name", ["cornell_2x2x2", "cornell_3x3x3", "field_8"])
def test_probe_update_headers_match_oracle(name, variant):
    cfg = CFG[name]
    sc = util.oracle_scene(cfg)
    rx, ry = cfg["tile"]
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))
    # field_8: 131072 rays; a 1/4 slab through the middle keeps the CPU suite short
    k0, k1 = (0, sc.num_rays) if sc.num_rays <= 4096 else (sc.num_rays // 2 - 16384, sc.num_rays // 2 + 16384)
    want = oracle.probe_update(sc, rays, k0, k1)
    alb, f32, lk = _sim_probe_update(sc, rays, variant, k0, k1)
    util.assert_lookups(lk[k0:k1], want[3][k0:k1], variant, name)
    assert np.array_equal(f32.view(np.uint32), want[2].view(np.uint32))
    assert np.array_equal(alb, want[0])


def test_wavefront_on_axis_parallel_and_degenerate_rays():
    """Directions with zero / tiny / NaN components take the literal two-division march
    (WF_MARCH_SLOW); origins on cell boundaries exercise fract == 0."""
    cfg = CFG["cornell_3x3x3"]
    sc = util.oracle_scene(cfg)
    dirs = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 0), (0, 1e-30, 1),
            (1e-25, 1, 1e-25), (0.6, 0.0, 0.8), (1, 1, 1), (-1, 2, -3), (0, 0, 0), (np.nan, 1, 0)]
    origins = [(0.0, 0.0, 15.0), (0.5, -0.5, 12.0), (-3.0, 2.0, 20.0), (1e-30, 0.0, 15.0)]
    n = len(dirs) * len(origins)
    rays = np.zeros((sc.num_rays, 12), dtype=np.float32)
    i = 0
    for o in origins:
        for d in dirs:
            rays[i, 0:3] = o
            rays[i, 4:7] = d
            rays[i, 8:11] = (i // 64, i % 8, (i % 64) // 8)
            i += 1
    with np.errstate(all="ignore"):
        want = oracle.probe_update(sc, rays, 0, n)
        for variant in (0, 1, 2):
            alb, f32, lk = _sim_probe_update(sc, rays, variant, 0, n)
            util.assert_lookups(lk[:n], want[3][:n], variant)
            assert np.array_equal(f32.view(np.uint32), want[2].view(np.uint32))
            assert np.array_equal(alb, want[0])


def test_frame_headers_match_oracle():
    cfg = util.small(CFG["cornell_3x3x3"], screen=(96, 96))
    sc = util.oracle_scene(cfg)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8, reseed=True))
    alb, *_ = oracle.probe_update(sc, rays)
    cam = util.camera_block(cfg)
    want, want_f32, want_lk = oracle.render_frame(sc, cam, alb)
    hs = util.hostsim()
    frame = np.zeros((96, 96), dtype=np.uint32)
    f32 = np.zeros((96, 96, 4), dtype=np.float32)
    lk = np.zeros((96, 96), dtype=np.uint32)
    hs.sim_render_frame(C.byref(sc.p), cam.ctypes.data, alb.ctypes.data, None, frame.ctypes.data, f32.ctypes.data, lk.ctypes.data)
    assert np.array_equal(lk, want_lk)
    assert np.array_equal(f32.view(np.uint32), want_f32.view(np.uint32))
    assert np.array_equal(frame, want)


def test_literal_colour_mode_matches_oracle_on_the_textured_cave():
    """Colour mode 1 (csrc/ddgi_texture.cuh: worley / fbm / dots / hash textures of
    intersection.glsl:872-1047) on the baked cave against the oracle's literal procedural mode,
    which tests/test_golden_reference.py pins to the reference's own shaders."""
    g = np.load(__import__("os").path.join(util.ROOT, "tests", "golden", "cave_3x3x3.npz"))
    vox = oracle.bake_scene(0, (128, 128, 128), (-64, -64, -64))
    sc = oracle.Scene(probe_count=(3, 3, 3), side_length=7, field_origin=(0.0, 0.0, 0.0), rx=8, lights=oracle.default_lights(0),
                      scene=0, voxels=vox, vorg=(-64, -64, -64), literal_colors=True, screen=tuple(int(v) for v in g["screen"]))
    rays = g["rays"]
    for variant in (0, 1, 2):
        alb, f32, lk = _sim_probe_update(sc, rays, variant)
        util.assert_lookups(lk, g["lookups"], variant)
        assert np.array_equal(f32.view(np.uint32), g["albedo_f32"].view(np.uint32))
        assert np.array_equal(alb, g["albedo"])
    hs = util.hostsim()
    w, h = sc.p.screen_width, sc.p.screen_height
    frame = np.zeros((h, w), dtype=np.uint32)
    f32 = np.zeros((h, w, 4), dtype=np.float32)
    lk = np.zeros((h, w), dtype=np.uint32)
    cam = np.ascontiguousarray(g["cam"])
    tex = np.ascontiguousarray(g["albedo"])
    hs.sim_render_frame(C.byref(sc.p), cam.ctypes.data, tex.ctypes.data, None, frame.ctypes.data, f32.ctypes.data, lk.ctypes.data)
    assert np.array_equal(lk, g["frame_lookups"])
    assert np.array_equal(f32.view(np.uint32), g["frame_f32"].view(np.uint32))
    assert np.array_equal(frame, g["frame"])


def test_hysteresis_blend_in_the_engine_headers():
    """blend_hysteresis (csrc/ddgi_trace.cuh) against the restored reference blend (golden)."""
    g = np.load(__import__("os").path.join(util.ROOT, "tests", "golden", "cornell_2x2x2.npz"))
    cfg = CFG["cornell_2x2x2"]
    sc = util.oracle_scene(cfg)
    sc.p.blend_mode = 1
    sc.p.hysteresis = float(g["hysteresis"])
    hs = util.hostsim()
    W, H = sc.tex_size
    rays = np.ascontiguousarray(g["rays"])
    for variant in (0, 1, 2):
        alb = np.zeros((H, W), dtype=np.uint32)
        for want in g["albedo_hysteresis"]:
            hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, 0, sc.num_rays, variant, alb.ctypes.data, None, None, None)
            assert np.array_equal(alb, want)
