"""What the reference's two shaders carry besides the default DDGI frame, each held to the
reference's own text:

  * the debug integrators, render_mode 1-5 (compute_pass.comp:58-87, integrators.glsl:110-271)
    and the probe markers of "Visualize Probes" (integrators.glsl:45-67, intersection.glsl:314-392)
    — live reference code, run unmodified;
  * the lines the reference has commented out, restored by removing the comment markers and
    nothing else (oracle/ref_glsl/build_ref.py RESTORE): `weight *= chebyshevWeight;`
    (intersection.glsl:1382) and `update_lights();` (probe_pass.comp:254, compute_pass.comp:174).

tests/golden/modes_*.npz are outputs of those builds of the transpiled shaders
(tests/golden/make_golden.py).  CPU tier: the oracle and the engine's headers (tests/hostsim)
against them; GPU tier: the CUDA engine through the C-ABI.  One extension has no reference
text: distance mode DDGI_DISTANCE_MOMENTS (the reference stores vec2(0)), checked engine vs oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest

import ddgi_b200
import util
from oracle import oracle, ref

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["modes_cornell_3x3x3", "modes_cave_3x3x3"]
MODES = (1, 2, 3, 4, 5, 9)
capi = ddgi_b200.capi


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def field(g):
    return dict(probe_count=tuple(int(v) for v in g["probe_count"]), side_length=int(g["side_length"]),
                field_origin=tuple(float(v) for v in g["field_origin"]))


def proc_scene(g, lights=None, **kw):
    """The oracle in literal procedural mode: the reference's own scene functions."""
    scene = int(g["scene"])
    return oracle.Scene(rx=int(g["s"]), lights=oracle.default_lights(scene) if lights is None else lights, scene=scene,
                        procedural=True, literal_colors=True, screen=tuple(int(v) for v in g["screen"]), **field(g), **kw)


def same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


# ------------------------------------------------------------------ CPU tier: oracle vs the reference shaders
@pytest.mark.parametrize("name", NAMES)
def test_oracle_debug_integrators_and_markers(name):
    g = load(name)
    for mode in MODES:
        sc = proc_scene(g, render_mode=mode)
        frame, f32, lk = oracle.render_frame(sc, g["cam"], g["albedo"])
        assert same(f32, g[f"frame_f32_mode{mode}"]), f"render_mode {mode}"
        assert np.array_equal(frame, g[f"frame_mode{mode}"])
        assert np.array_equal(lk, g[f"frame_lookups_mode{mode}"])
    for mode in (0, 2):
        sc = proc_scene(g, render_mode=mode, visualize_probes=True)
        frame, f32, _ = oracle.render_frame(sc, g["cam"], g["albedo"])
        assert same(f32, g[f"frame_f32_markers_mode{mode}"])
        assert np.array_equal(frame, g[f"frame_markers_mode{mode}"])
        assert (frame == 0xFFFFFF00).any(), "no marker pixel (cyan) in the fixture view"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_chebyshev_weight(name):
    g = load(name)
    sc = proc_scene(g, chebyshev=True)
    for tag, dist in (("zero", np.zeros_like(g["albedo"])), ("random", g["distances_random"])):
        frame, f32, _ = oracle.render_frame(sc, g["cam"], g["albedo"], tex_distances=dist)
        assert same(f32, g[f"frame_f32_chebyshev_{tag}"])
        assert np.array_equal(frame, g[f"frame_chebyshev_{tag}"])
    assert not np.array_equal(g["frame_chebyshev_zero"], g["frame_chebyshev_random"])  # the distance image does matter


@pytest.mark.parametrize("name", NAMES)
def test_oracle_update_lights(name):
    g = load(name)
    scene, t = int(g["scene"]), float(g["lights_time"])
    sc = proc_scene(g, lights=oracle.update_lights(scene, oracle.default_lights(scene), t))
    alb, _, f32, lk, _ = oracle.probe_update(sc, g["rays"])
    assert np.array_equal(lk, g["lookups_lights"])
    assert same(f32, g["albedo_f32_lights"]) and np.array_equal(alb, g["albedo_lights"])
    frame, ff32, _ = oracle.render_frame(sc, g["cam"], alb)
    assert same(ff32, g["frame_f32_lights"]) and np.array_equal(frame, g["frame_lights"])
    assert not np.array_equal(alb, g["albedo"])  # the lights did move


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_oracle_modes_match_live_reference_on_the_house():
    """Scene 2 (two lights) is in no modes fixture: compare live."""
    scene, pc, side, org, s, screen = 2, (3, 1, 3), 9, (0.0, 0.0, 0.0), 4, (48, 32)
    kw = dict(probe_count=pc, side_length=side, field_origin=org)
    lights = oracle.update_lights(scene, oracle.default_lights(scene), 4000.0)
    sc = oracle.Scene(rx=s, lights=lights, scene=scene, procedural=True, literal_colors=True, screen=screen, **kw)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
    want = ref.probe_pass(scene=scene, s=s, rays=rays, animate_lights_time=4000.0, **kw)
    got = oracle.probe_update(sc, rays)
    assert np.array_equal(got[0], want[0]) and same(got[2], want[2]) and np.array_equal(got[3], want[3])
    cam = ddgi_b200.Camera(screen[0] / float(screen[1]), (0.0, 0.0, -10.0), (0.0, 0.0, 0.0)).get_data()
    for mode in (1, 4, 5):
        sc2 = oracle.Scene(rx=s, lights=oracle.default_lights(scene), scene=scene, procedural=True, literal_colors=True,
                           screen=screen, render_mode=mode, visualize_probes=True, **kw)
        a = oracle.render_frame(sc2, cam, got[0])
        b = ref.compute_pass(scene=scene, s=s, screen=screen, cam=cam, tex_albedo=got[0], render_mode=mode, visualize_probes=True, **kw)
        assert np.array_equal(a[0], b[0]) and same(a[1], b[1])


# ------------------------------------------------------------------ CPU tier: the engine's headers (hostsim)
def stored_scene(g, **kw):
    """The engine's scene model for a fixture: stored voxels; Cornell's colours are flat, the cave
    needs the literal colour mode (both are then exactly the procedural scene)."""
    scene = int(g["scene"])
    if scene == 1:
        vox, vorg = oracle.bake_scene(1, (32, 32, 32), (-15, -15, 0)), (-15, -15, 0)
    else:
        vox, vorg = oracle.bake_scene(0, (128, 128, 128), (-64, -64, -64)), (-64, -64, -64)
    lights = kw.pop("lights", None)
    return oracle.Scene(rx=int(g["s"]), lights=oracle.default_lights(scene) if lights is None else lights, scene=scene, voxels=vox,
                        vorg=vorg, literal_colors=(scene == 0), screen=tuple(int(v) for v in g["screen"]), **field(g), **kw)


def sim_frame(sc, cam, tex, dist=None):
    hs = util.hostsim()
    w, h = sc.p.screen_width, sc.p.screen_height
    frame = np.zeros((h, w), dtype=np.uint32)
    f32 = np.zeros((h, w, 4), dtype=np.float32)
    cam = np.ascontiguousarray(cam)
    tex = np.ascontiguousarray(tex)
    dist = np.zeros_like(tex) if dist is None else np.ascontiguousarray(dist)
    hs.sim_render_frame(C.byref(sc.p), cam.ctypes.data, tex.ctypes.data, dist.ctypes.data, frame.ctypes.data, f32.ctypes.data, None)
    return frame, f32


@pytest.mark.parametrize("name", NAMES)
def test_engine_headers_modes(name):
    g = load(name)
    for mode in MODES:
        frame, f32 = sim_frame(stored_scene(g, render_mode=mode), g["cam"], g["albedo"])
        assert same(f32, g[f"frame_f32_mode{mode}"]), f"render_mode {mode}"
        assert np.array_equal(frame, g[f"frame_mode{mode}"])
    for mode in (0, 2):
        frame, f32 = sim_frame(stored_scene(g, render_mode=mode, visualize_probes=True), g["cam"], g["albedo"])
        assert same(f32, g[f"frame_f32_markers_mode{mode}"])
    for tag, dist in (("zero", None), ("random", g["distances_random"])):
        frame, f32 = sim_frame(stored_scene(g, chebyshev=True), g["cam"], g["albedo"], dist)
        assert same(f32, g[f"frame_f32_chebyshev_{tag}"])
        assert np.array_equal(frame, g[f"frame_chebyshev_{tag}"])


def test_engine_headers_distance_moments_match_oracle():
    """DDGI_DISTANCE_MOMENTS (no reference text): both kernel variants' first-hit t against the oracle."""
    g = load("modes_cornell_3x3x3")
    for scale in (1.0, 19.0):
        sc = stored_scene(g, distance_scale=scale)
        want = oracle.probe_update(sc, g["rays"])
        assert (want[1] != 0).any()
        hs = util.hostsim()
        rays = np.ascontiguousarray(g["rays"])
        for variant in (0, 1, 2):
            alb = np.zeros_like(want[0])
            dist = np.zeros_like(want[0])
            hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, 0, sc.num_rays, variant, alb.ctypes.data, None, None, dist.ctypes.data)
            assert np.array_equal(alb, want[0]) and np.array_equal(dist, want[1])
        # and the Chebyshev sample over those moments
        sc2 = stored_scene(g, distance_scale=scale, chebyshev=True)
        a = oracle.render_frame(sc2, g["cam"], want[0], tex_distances=want[1])
        frame, f32 = sim_frame(sc2, g["cam"], want[0], want[1])
        assert same(f32, a[1]) and np.array_equal(frame, a[0])


# ------------------------------------------------------------------ GPU tier: the CUDA engine through the C-ABI
def engine(g):
    scene = int(g["scene"])
    w, h = (int(v) for v in g["screen"])
    r = ddgi_b200.RVPT(w, h)
    r.set_debug(True)
    r.render_settings.scene = scene
    r.ir.probe_count[:] = tuple(int(v) for v in g["probe_count"])
    r.ir.side_length = int(g["side_length"])
    r.ir.sqrt_rays_per_probe = int(g["s"])
    r.ir.field_origin[:] = tuple(float(v) for v in g["field_origin"])
    if scene == 1:
        r.bake_scene((32, 32, 32), (-15, -15, 0), scene=1)
        r.scene_camera = ddgi_b200.Camera(w / float(h), (0.0, 0.0, -5.0), (0.0, 0.0, 0.0))
    else:
        r.bake_scene((128, 128, 128), (-64, -64, -64), scene=0)
        r.set_color_mode(capi.COLOR_LITERAL)
        r.scene_camera = ddgi_b200.Camera(w / float(h), (1.5, 2.0, -2.0), (-38.0, 36.0, 0.0))
    r.set_probe_rays(g["rays"])
    assert same(r.scene_camera.get_data(), g["cam"])
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_debug_integrators_and_markers(name):
    g = load(name)
    with engine(g) as r:
        r.update(advance_time=False)
        r.probe_update()
        r.sync()
        assert np.array_equal(r.read_probe_texture(0), g["albedo"])
        for mode in MODES:
            r.render_settings.render_mode = mode
            r.update(advance_time=False)
            r.render_frame()
            r.sync()
            assert same(r.read_frame(capi.FMT_F32), g[f"frame_f32_mode{mode}"]), f"render_mode {mode}"
            assert np.array_equal(r.read_frame(), g[f"frame_mode{mode}"])
            w, h = (int(v) for v in g["screen"])
            assert np.array_equal(r.read_lookup_counts(1).reshape(h, w), g[f"frame_lookups_mode{mode}"])
        r.render_settings.visualize_probes = 1
        for mode in (0, 2):
            r.render_settings.render_mode = mode
            r.update(advance_time=False)
            r.render_frame()
            r.sync()
            assert same(r.read_frame(capi.FMT_F32), g[f"frame_f32_markers_mode{mode}"])
            assert np.array_equal(r.read_frame(), g[f"frame_markers_mode{mode}"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_chebyshev_weight(name):
    g = load(name)
    with engine(g) as r:
        r.update(advance_time=False)
        r.set_weight_mode(capi.WEIGHT_CHEBYSHEV)
        r.write_probe_texture(g["albedo"], 0)
        for tag, dist in (("zero", np.zeros_like(g["albedo"])), ("random", g["distances_random"])):
            r.write_probe_texture(dist, 1)
            r.render_frame()
            r.sync()
            assert same(r.read_frame(capi.FMT_F32), g[f"frame_f32_chebyshev_{tag}"])
            assert np.array_equal(r.read_frame(), g[f"frame_chebyshev_{tag}"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_cuda_update_lights(name, variant):
    g = load(name)
    with engine(g) as r:
        r.set_kernel_variant(variant)
        r.animate_lights = True
        r.render_settings.time = float(g["lights_time"])
        r.update(advance_time=False)
        r.draw()
        r.sync()
        util.assert_lookups(r.read_lookup_counts(0), g["lookups_lights"], variant, name)
        assert same(r.read_probe_texture(0, capi.FMT_F32), g["albedo_f32_lights"])
        assert np.array_equal(r.read_probe_texture(0), g["albedo_lights"])
        assert same(r.read_frame(capi.FMT_F32), g["frame_f32_lights"])
        assert np.array_equal(r.read_frame(), g["frame_lights"])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_cuda_distance_moments_and_checkpoint(variant, tmp_path):
    g = load("modes_cornell_3x3x3")
    scale = 19.0
    sc = stored_scene(g, distance_scale=scale)
    want = oracle.probe_update(sc, g["rays"])
    sc2 = stored_scene(g, distance_scale=scale, chebyshev=True)
    want_frame = oracle.render_frame(sc2, g["cam"], want[0], tex_distances=want[1])
    with engine(g) as r:
        r.set_kernel_variant(variant)
        r.set_distance_mode(capi.DISTANCE_MOMENTS, scale)
        r.set_weight_mode(capi.WEIGHT_CHEBYSHEV)
        r.update(advance_time=False)
        r.draw()
        r.sync()
        assert np.array_equal(r.read_probe_texture(0), want[0])
        assert np.array_equal(r.read_probe_texture(1), want[1]) and (want[1] != 0).any()
        assert same(r.read_frame(capi.FMT_F32), want_frame[1])
        # checkpoint: both planes + time to disk, wiped, restored
        path = str(tmp_path / "probe.ckpt")
        r.render_settings.time = 12.0
        r.save_checkpoint(path)
        r.write_probe_texture(np.zeros_like(want[0]), 0)
        r.write_probe_texture(np.zeros_like(want[0]), 1)
        r.render_settings.time = 0.0
        r.load_checkpoint(path)
        assert r.render_settings.time == 12.0
        assert np.array_equal(r.read_probe_texture(0), want[0]) and np.array_equal(r.read_probe_texture(1), want[1])
        r.render_frame()
        r.sync()
        assert same(r.read_frame(capi.FMT_F32), want_frame[1])


# ------------------------------------------------------------------ dynamic scenes (no reference text: its scene is compiled in)
def test_update_lights_host_function_matches_oracle():
    """ddgi_update_lights is a host-side function of the C-ABI: callable without a GPU."""
    lib = capi.load()
    for scene in (0, 1, 2):
        base = oracle.default_lights(scene) if scene != 0 else oracle.cave_lights4(0.0)
        if scene == 0:  # the 4-light table: exercises i >= 1
            arr0 = (oracle.OrcLight * 8)()
            oracle.load().orc_cave_lights4(arr0)
            base = [arr0[i] for i in range(4)]
        n = len(base)
        arr = (capi.Light * 8)()
        for i, l in enumerate(base):
            arr[i].intensity = l.intensity
            arr[i].col[:] = list(l.col)
            arr[i].pos[:] = list(l.pos)
        out = (capi.Light * 8)()
        for t in (0.0, 2.0, 74.0, 12345.0):
            assert lib.ddgi_update_lights(scene, t, arr, n, out) == 0
            want = oracle.update_lights(scene, base, t)
            for i in range(n):
                assert same(np.array(list(out[i].pos), dtype=np.float32), np.array(list(want[i].pos), dtype=np.float32))
                assert out[i].intensity == want[i].intensity
    assert lib.ddgi_update_lights(3, 0.0, arr, 1, out) == capi.E_INVALID


@pytest.mark.gpu
def test_cuda_voxel_edit_matches_reupload_and_oracle(tmp_path):
    """ddgi_edit_voxels: a pillar placed into (and a hole cut out of) the Cornell box frame by
    frame; the occupancy bricks are rebuilt only where touched, the result must equal a full
    re-upload and the oracle on the edited field."""
    cfg = util.small(util.configs.CONFIGS["cornell_3x3x3"], screen=(64, 64))
    vox, vorg = util.oracle_voxels(cfg)
    vox = vox.copy()
    rx, ry = cfg["tile"]
    edits = [((-2, -9, 9), np.full((7, 13, 3), 4, dtype=np.uint8)),     # [z, y, x] blue pillar, odd sizes / offsets
             ((-10, -3, 12), np.zeros((5, 5, 1), dtype=np.uint8)),       # a window in the red wall
             ((16, 16, 31), np.full((1, 1, 1), 3, dtype=np.uint8))]      # last cell of the grid
    with ddgi_b200.RVPT(64, 64) as r, ddgi_b200.RVPT(64, 64) as fresh:
        r.set_debug(True)
        util.configs.apply(r, cfg)
        r.generate_probe_rays(reseed=True)
        r.update(advance_time=False)
        r.probe_update()  # calibrates the schedule on the unedited scene
        for origin, box in edits:
            r.edit_voxels(box, origin)
            z0, y0, x0 = origin[2] - vorg[2], origin[1] - vorg[1], origin[0] - vorg[0]
            vox[z0:z0 + box.shape[0], y0:y0 + box.shape[1], x0:x0 + box.shape[2]] = box
            assert np.array_equal(r.read_voxels(cfg["voxels"][1]), vox)
            sc = util.oracle_scene(cfg, voxels=vox)
            rays = oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))
            want = oracle.probe_update(sc, rays)
            for variant in (0, 1, 2):
                r.set_kernel_variant(variant)
                r.draw()
                r.sync()
                util.assert_lookups(r.read_lookup_counts(0), want[3], variant)
                assert np.array_equal(r.read_probe_texture(0), want[0])
            want_frame = oracle.render_frame(sc, util.camera_block(cfg), want[0])
            assert np.array_equal(r.read_frame(), want_frame[0])
        # out-of-field edits are refused, not clipped
        with pytest.raises(ddgi_b200.DDGIError) as e:
            r.edit_voxels(np.zeros((2, 2, 2), dtype=np.uint8), (16, 16, 31))
        assert e.value.code == capi.E_INVALID
        # voxel file round trip into a fresh context = the same texture
        path = str(tmp_path / "scene.vox")
        r.save_voxels(path, cfg["voxels"][1], vorg)
        util.configs.apply(fresh, cfg)
        dims, origin = fresh.load_voxels(path)
        assert dims == tuple(cfg["voxels"][1]) and origin == tuple(vorg)
        fresh.generate_probe_rays(reseed=True)
        fresh.update(advance_time=False)
        fresh.probe_update()
        fresh.sync()
        assert np.array_equal(fresh.read_probe_texture(0), r.read_probe_texture(0))
