"""The oracle is pinned against the REFERENCE'S OWN SHADERS.

tests/golden/*.npz are outputs of probe_pass.comp / compute_pass.comp themselves, transpiled
where they lie in /root/reference and run on the CPU (oracle/ref_glsl/build_ref.py,
tests/golden/make_golden.py).  The oracle restatement (oracle/ddgi_oracle.c, literal
procedural mode) must reproduce them bit for bit: RGBA8 bytes, the fp32 values handed to
imageStore, and the number of getBlockAt calls per invocation.  When oracle/_ref is built
(always in the build container, prebuilt on the GPU box) the same comparison also runs live
against the transpiled shaders on inputs the fixtures do not hold.
"""
import glob
import os

import numpy as np
import pytest

import ddgi_b200
import util
from oracle import oracle, ref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith(("modes_", "full_")))


def _scene(g, screen=None):
    scene = int(g["scene"])
    s = int(g["s"])
    return oracle.Scene(probe_count=tuple(int(v) for v in g["probe_count"]), side_length=int(g["side_length"]),
                        field_origin=tuple(float(v) for v in g["field_origin"]), rx=s, lights=oracle.default_lights(scene),
                        scene=scene, procedural=True, literal_colors=True,
                        screen=tuple(int(v) for v in (g["screen"] if screen is None else screen)))


def test_fixtures_exist():
    assert len(GOLDEN) >= 4, "tests/golden/*.npz are committed fixtures"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_the_reference_shaders(path):
    g = np.load(path)
    sc = _scene(g)
    s = int(g["s"])
    # the ProbeRay list: the reference builds it on the host (rvpt.cpp:1177-1224)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
    assert np.array_equal(rays.view(np.uint32), g["rays"].view(np.uint32))
    alb, dist, f32, lk, _ = oracle.probe_update(sc, rays)
    assert np.array_equal(lk, g["lookups"]), "getBlockAt counts differ: a ray took another discrete path"
    assert np.array_equal(f32.view(np.uint32), g["albedo_f32"].view(np.uint32))
    assert np.array_equal(alb, g["albedo"])
    assert np.array_equal(dist, g["distances"]) and (dist == 0).all()
    frame, frame_f32, frame_lk = oracle.render_frame(sc, g["cam"], g["albedo"])
    assert np.array_equal(frame_lk, g["frame_lookups"])
    assert np.array_equal(frame_f32.view(np.uint32), g["frame_f32"].view(np.uint32))
    assert np.array_equal(frame, g["frame"])


def test_oracle_reproduces_the_reference_shaders_on_cave_128_at_full_size():
    """BASELINE.json configs[2] (16^3 probes x 256 rays, 1080p) with the reference's procedural scene and textures: all
    1 048 576 probe rays and all 2 073 600 pixels of the oracle against the reference's own shaders run at full size
    (tests/golden/full_cave_128.npz: CRC-32 / SHA-256 of the whole outputs, 94 probe tiles, 64 frame rows)."""
    import hashlib
    import zlib

    g = np.load(os.path.join(HERE, "golden", "full_cave_128.npz"))
    sc = _scene(dict(scene=0, s=g["s"], probe_count=g["probe_count"], side_length=g["side_length"], field_origin=g["field_origin"],
                     screen=g["screen"]))
    s = int(g["s"])
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
    alb, dist, f32, lk, _ = oracle.probe_update(sc, rays)
    X, Y, Z = (int(v) for v in g["probe_count"])
    tiles = alb.reshape(Y, s, X * Z, s).transpose(0, 2, 1, 3).reshape(X * Y * Z, s, s)
    tiles_f32 = f32.reshape(Y, s, X * Z, s, 4).transpose(0, 2, 1, 3, 4).reshape(X * Y * Z, s, s, 4)
    assert np.array_equal(tiles[g["probes"]], g["tiles"])
    assert np.array_equal(tiles_f32[g["probes"]].view(np.uint32), g["tiles_f32"].view(np.uint32))
    assert np.array_equal(lk.reshape(X * Y * Z, s * s)[g["probes"]], g["tile_lookups"])
    assert int(lk.sum(dtype=np.uint64)) == int(g["lookups_sum"])
    assert zlib.crc32(alb.tobytes()) == int(g["albedo_crc32"]) and hashlib.sha256(alb.tobytes()).hexdigest() == str(g["albedo_sha256"])
    assert (dist == 0).all()
    frame, _, flk = oracle.render_frame(sc, g["cam"], alb)
    b0, b1 = (int(v) for v in g["band"])
    assert np.array_equal(frame[b0:b1], g["frame_band"])
    assert zlib.crc32(frame.tobytes()) == int(g["frame_crc32"]) and hashlib.sha256(frame.tobytes()).hexdigest() == str(g["frame_sha256"])
    assert int(flk.sum(dtype=np.uint64)) == int(g["frame_lookups_sum"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_hysteresis_mode_reproduces_the_restored_reference_blend(path):
    """blend_mode 1 against probe_pass.comp built with the comment markers around its own
    hysteresis blend (:298-299) removed: three successive frames from a zero texture."""
    g = np.load(path)
    scene, s = int(g["scene"]), int(g["s"])
    sc = oracle.Scene(probe_count=tuple(int(v) for v in g["probe_count"]), side_length=int(g["side_length"]),
                      field_origin=tuple(float(v) for v in g["field_origin"]), rx=s, lights=oracle.default_lights(scene),
                      scene=scene, procedural=True, literal_colors=True, hysteresis=float(g["hysteresis"]))
    tex = np.zeros_like(g["albedo"])
    for want in g["albedo_hysteresis"]:
        oracle.probe_update(sc, g["rays"], tex=tex)
        assert np.array_equal(tex, want)
    assert not np.array_equal(g["albedo_hysteresis"][0], g["albedo_hysteresis"][2])  # the texture does carry state


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference; prebuilt .so travels to the GPU box)")
@pytest.mark.parametrize("scene,pc,side,org,s,cam_o,cam_r", [
    (1, (2, 3, 2), 9, (1.0, -1.0, 14.0), 6, (2.0, 1.0, -4.0), (-10.0, 8.0, 0.0)),   # even/odd mix, s not a power of two
    (0, (2, 2, 2), 11, (1.4, 0.0, 1.0), 4, (1.5, 2.0, -2.0), (-38.0, 36.0, 0.0)),   # the reference's default field origin
    (0, (1, 1, 1), 5, (0.0, -14.0, 0.0), 10, (0.0, -12.0, 0.0), (20.0, 200.0, 0.0)),  # probe just above the fbm floor band
])
def test_oracle_matches_live_reference_shaders(scene, pc, side, org, s, cam_o, cam_r):
    screen = (64, 48)
    sc = oracle.Scene(probe_count=pc, side_length=side, field_origin=org, rx=s, lights=oracle.default_lights(scene),
                      scene=scene, procedural=True, literal_colors=True, screen=screen)
    rng = np.random.default_rng(scene * 7 + s)
    samples = rng.normal(size=(s * s, 3)).astype(np.float32)  # any directions: the shader normalises
    rays = oracle.generate_probe_rays(sc, samples)
    want = ref.probe_pass(scene=scene, probe_count=pc, side_length=side, field_origin=org, s=s, rays=rays)
    alb, dist, f32, lk, _ = oracle.probe_update(sc, rays)
    assert np.array_equal(lk, want[3])
    assert np.array_equal(f32.view(np.uint32), want[2].view(np.uint32))
    assert np.array_equal(alb, want[0]) and np.array_equal(dist, want[1])
    cam = ddgi_b200.Camera(screen[0] / float(screen[1]), cam_o, cam_r).get_data()
    gf = ref.compute_pass(scene=scene, probe_count=pc, side_length=side, field_origin=org, s=s, screen=screen, cam=cam,
                          tex_albedo=alb)
    frame, frame_f32, frame_lk = oracle.render_frame(sc, cam, alb)
    assert np.array_equal(frame_lk, gf[2])
    assert np.array_equal(frame_f32.view(np.uint32), gf[1].view(np.uint32))
    assert np.array_equal(frame, gf[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell_2x2x2", "cornell_3x3x3"])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_cuda_engine_reproduces_the_reference_shaders_on_cornell(name, variant):
    """The CUDA path against the reference shader outputs directly (no oracle in between).
    Cornell's block types 2-5 have flat colours (intersection.glsl:908-919) and the whole
    scene fits the baked 32^3 box, so the stored-voxel engine must match the procedural
    reference exactly."""
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    cfg = util.small(util.configs.CONFIGS[name], screen=tuple(int(v) for v in g["screen"]))
    with ddgi_b200.RVPT(*cfg["screen"]) as r:
        r.set_debug(True)
        util.configs.apply(r, cfg)
        r.set_probe_rays(g["rays"])          # literal storage-buffer mode: the fixture's ProbeRay list
        r.set_kernel_variant(variant)
        r.update(advance_time=False)
        r.draw()
        r.sync()
        util.assert_lookups(r.read_lookup_counts(0), g["lookups"], variant)
        assert np.array_equal(r.read_probe_texture(0, ddgi_b200.capi.FMT_F32).view(np.uint32), g["albedo_f32"].view(np.uint32))
        assert np.array_equal(r.read_probe_texture(0), g["albedo"])
        assert np.array_equal(r.read_probe_texture(1), g["distances"])
        assert np.array_equal(r.read_frame(), g["frame"])
        assert np.array_equal(r.read_frame(ddgi_b200.capi.FMT_F32).view(np.uint32), g["frame_f32"].view(np.uint32))
        # generated mode (no 48 B/ray buffer) gives the same texture
        r.generate_probe_rays(reseed=True)
        r.probe_update()
        r.sync()
        assert np.array_equal(r.read_probe_texture(0), g["albedo"])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_cuda_engine_reproduces_the_reference_shaders_on_the_textured_cave(variant):
    """Colour mode DDGI_COLOR_LITERAL: the reference's procedural cave textures on the device.
    The cave is baked over [-64,64)^3, which holds every probe and every surface a ray from
    them can reach, so the stored-voxel engine must match the procedural reference exactly —
    probe texture and final frame, bytes, fp32 bits and getBlockAt counts."""
    g = np.load(os.path.join(HERE, "golden", "cave_3x3x3.npz"))
    w, h = (int(v) for v in g["screen"])
    with ddgi_b200.RVPT(w, h) as r:
        r.set_debug(True)
        r.render_settings.scene = 0
        r.ir.probe_count[:] = (3, 3, 3)
        r.ir.side_length = 7
        r.ir.sqrt_rays_per_probe = 8
        r.ir.field_origin[:] = (0.0, 0.0, 0.0)
        r.scene_camera = ddgi_b200.Camera(w / float(h), (1.5, 2.0, -2.0), (-38.0, 36.0, 0.0))
        r.bake_scene((128, 128, 128), (-64, -64, -64), scene=0)
        r.set_color_mode(ddgi_b200.capi.COLOR_LITERAL)
        r.set_kernel_variant(variant)
        r.set_probe_rays(g["rays"])
        r.update(advance_time=False)
        assert np.array_equal(r.scene_camera.get_data().view(np.uint32), g["cam"].view(np.uint32))
        r.draw()
        r.sync()
        util.assert_lookups(r.read_lookup_counts(0), g["lookups"], variant)
        assert np.array_equal(r.read_probe_texture(0, ddgi_b200.capi.FMT_F32).view(np.uint32), g["albedo_f32"].view(np.uint32))
        assert np.array_equal(r.read_probe_texture(0), g["albedo"])
        assert np.array_equal(r.read_lookup_counts(1).reshape(h, w), g["frame_lookups"])
        assert np.array_equal(r.read_frame(ddgi_b200.capi.FMT_F32).view(np.uint32), g["frame_f32"].view(np.uint32))
        assert np.array_equal(r.read_frame(), g["frame"])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_cuda_engine_hysteresis_mode_reproduces_the_restored_reference_blend(variant):
    """DDGI_BLEND_HYSTERESIS on Cornell against the reference shader with its own blend restored."""
    g = np.load(os.path.join(HERE, "golden", "cornell_3x3x3.npz"))
    cfg = util.small(util.configs.CONFIGS["cornell_3x3x3"], screen=(64, 64))
    with ddgi_b200.RVPT(64, 64) as r:
        util.configs.apply(r, cfg)
        r.ir.hysteresis = float(g["hysteresis"])
        r.set_blend_mode(ddgi_b200.capi.BLEND_HYSTERESIS)
        r.set_kernel_variant(variant)
        r.set_probe_rays(g["rays"])
        r.update(advance_time=False)
        r.write_probe_texture(np.zeros_like(g["albedo"]))
        for want in g["albedo_hysteresis"]:
            r.probe_update()
            r.sync()
            assert np.array_equal(r.read_probe_texture(0), want)
        # checkpoint / resume: frame 3 from a re-uploaded frame 2
        r.write_probe_texture(g["albedo_hysteresis"][1])
        r.probe_update()
        r.sync()
        assert np.array_equal(r.read_probe_texture(0), g["albedo_hysteresis"][2])
