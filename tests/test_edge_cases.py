"""Edge cases of the probe update and the pixel pass the regular scenes rarely reach, CPU tier
(the engine's headers on the host, tests/hostsim, both kernel variants, against the oracle) and
GPU tier (the CUDA engine through the C-ABI):

  * no light at all, and the maximum of 8 lights;
  * light spheres as the NEAREST hit of a bounce ray or of a shadow feeler (lights placed a few
    tenths of a unit from probe origins and in front of walls): intersection.glsl:1262-1279 and
    the type-2 branches of probe_pass.comp:186-207 / integrators.glsl:73,85;
  * 1 bounce and 64 bounces (the engine's maximum), odd ray-tile sizes, one ray per probe;
  * a probe field far outside the voxel box (every march runs its 125 steps through empty space)
    and one straddling the box boundary.
"""
import ctypes as C

import numpy as np
import pytest

import ddgi_b200
import util
from oracle import oracle

capi = ddgi_b200.capi
CFG = util.configs.CONFIGS


def light(intensity, col, pos):
    l = oracle.OrcLight()
    l.intensity = intensity
    l.col[:] = col
    l.pos[:] = pos
    return l


def lights_near_probes(cfg, count):
    """`count` lights, each 0.13-0.22 from a probe origin of the field (so some of that probe's rays
    hit the radius-0.1 sphere first) — the last one in front of the back wall instead."""
    X, Y, Z = cfg["probe_count"]
    side, org = cfg["side_length"], np.array(cfg["field_origin"], dtype=np.float64)
    rng = np.random.default_rng(count)
    out = []
    for i in range(count):
        idx = np.array([rng.integers(0, X), rng.integers(0, Y), rng.integers(0, Z)])
        pos = (idx - (np.array([X, Y, Z]) - 1) // 2) * side + org
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        pos = pos + d * rng.uniform(0.13, 0.22)
        out.append(light(float(rng.uniform(5, 40)), [float(v) for v in rng.uniform(0.2, 1.2, size=3)], [float(np.float32(v)) for v in pos]))
    if count >= 2:
        out[-1] = light(25.0, [1.0, 0.9, 0.8], [0.5, 0.25, 24.6])   # hugging the z = 25 wall of the Cornell box
    return out


def case_scene(case):
    cfg = dict(CFG["cornell_3x3x3"])
    cfg["screen"] = (64, 48)
    kw = dict(max_bounces=8)
    lights = oracle.default_lights(1)
    tile = (8, 8)
    if case == "no_lights":
        lights = []
    elif case == "eight_lights_near_probes":
        lights = lights_near_probes(cfg, 8)
    elif case == "three_lights_near_probes_5x7":
        lights = lights_near_probes(cfg, 3)
        tile = (5, 7)
    elif case == "one_bounce":
        kw["max_bounces"] = 1
    elif case == "sixty_four_bounces":
        kw["max_bounces"] = 64
        cfg["probe_count"] = (1, 1, 1)
    elif case == "one_ray_per_probe":
        tile = (1, 1)
        cfg["probe_count"] = (2, 1, 3)
    elif case == "field_outside_the_box":
        cfg["field_origin"] = (400.0, -300.0, 15.0)
    elif case == "field_straddling_the_box":
        cfg["field_origin"] = (9.0, 9.0, 28.0)
    vox, vorg = util.oracle_voxels(CFG["cornell_3x3x3"])
    sc = oracle.Scene(probe_count=cfg["probe_count"], side_length=cfg["side_length"], field_origin=cfg["field_origin"], rx=tile[0],
                      ry=tile[1], lights=lights, scene=1, voxels=vox, vorg=vorg, screen=cfg["screen"], **kw)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(tile[0], tile[1], reseed=True))
    return cfg, sc, rays, lights, tile, kw["max_bounces"]


CASES = ["no_lights", "eight_lights_near_probes", "three_lights_near_probes_5x7", "one_bounce", "sixty_four_bounces",
         "one_ray_per_probe", "field_outside_the_box", "field_straddling_the_box"]


def light_hits(sc, rays):
    """How many probe rays have a light sphere as their first hit (oracle's intersect_scene)."""
    n = 0
    out = np.zeros(13, dtype=np.float32)
    lib = oracle.load()
    for k in range(0, rays.shape[0], 1):
        o = np.ascontiguousarray(rays[k, 0:3])
        d = np.ascontiguousarray(rays[k, 4:7])
        lib.orc_intersect_scene(C.byref(sc.p), o.ctypes.data, d.ctypes.data, out.ctypes.data)
        n += int(out[0] == 1.0 and out[11] == 2.0)
    return n


@pytest.mark.parametrize("case", CASES)
def test_engine_headers_edge_cases(case):
    cfg, sc, rays, lights, tile, _ = case_scene(case)
    with np.errstate(all="ignore"):
        want = oracle.probe_update(sc, rays)
    if "near_probes" in case:
        assert light_hits(sc, rays) >= 3, "the case must put light spheres in front of some probe rays"
    if case == "field_outside_the_box":
        assert (want[3] == 125).all() and (want[0] == 0xFF000000).all()   # every ray: one march of 125 empty cells, black
    hs = util.hostsim()
    rays = np.ascontiguousarray(rays)
    for variant in (0, 1, 2):
        alb = np.zeros_like(want[0])
        f32 = np.zeros_like(want[2])
        lk = np.zeros_like(want[3])
        hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, 0, sc.num_rays, variant, alb.ctypes.data, f32.ctypes.data, lk.ctypes.data, None)
        util.assert_lookups(lk, want[3], variant, case)
        assert np.array_equal(f32.view(np.uint32), want[2].view(np.uint32)), f"{case}: fp32 texels, variant {variant}"
        assert np.array_equal(alb, want[0])
    # pixel pass: looking at the lights (emissive pixels) through the same scene
    cam = util.camera_block(cfg)
    f = oracle.render_frame(sc, cam, want[0])
    w, h = sc.p.screen_width, sc.p.screen_height
    frame, ff32, flk = np.zeros((h, w), dtype=np.uint32), np.zeros((h, w, 4), dtype=np.float32), np.zeros((h, w), dtype=np.uint32)
    hs.sim_render_frame(C.byref(sc.p), cam.ctypes.data, want[0].ctypes.data, None, frame.ctypes.data, ff32.ctypes.data, flk.ctypes.data)
    assert np.array_equal(flk, f[2])
    assert np.array_equal(ff32.view(np.uint32), f[1].view(np.uint32)) and np.array_equal(frame, f[0])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_edge_cases(case):
    cfg, sc, rays, lights, tile, bounces = case_scene(case)
    with np.errstate(all="ignore"):
        want = oracle.probe_update(sc, rays)
    f = oracle.render_frame(sc, util.camera_block(cfg), want[0])
    with ddgi_b200.RVPT(*cfg["screen"]) as r:
        r.set_debug(True)
        util.configs.apply(r, cfg)
        r.ray_tile = tile
        r.render_settings.max_bounces = bounces
        r.lights = [capi.Light(l.intensity, tuple(l.col), tuple(l.pos)) for l in lights]
        r.generate_probe_rays(reseed=True)
        for variant in (0, 1, 2):
            r.set_kernel_variant(variant)
            r.update(advance_time=False)
            r.draw()
            r.sync()
            util.assert_lookups(r.read_lookup_counts(0), want[3], variant, case)
            assert np.array_equal(r.read_probe_texture(0, capi.FMT_F32).view(np.uint32), want[2].view(np.uint32))
            assert np.array_equal(r.read_probe_texture(0), want[0])
            assert np.array_equal(r.read_frame(capi.FMT_F32).view(np.uint32), f[1].view(np.uint32))
            assert np.array_equal(r.read_frame(), f[0])
