"""Randomised parity, CPU tier: random voxel boxes (odd sizes, origins that are not multiples of the
4x4x2 brick, 5-40 % solid), random lights (some inside the box, some far away), random probe fields
and ray tiles — the engine's headers (tests/hostsim, both kernel variants) against the oracle, bit
for bit: texels, fp32 values and per-ray lookup counts, then one frame.  A GPU-tier twin runs the
same generator through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

import ddgi_b200
import util
from oracle import oracle

capi = ddgi_b200.capi


def random_case(seed):
    rng = np.random.default_rng(seed)
    dims = tuple(int(v) for v in rng.integers(5, 29, size=3))          # (dx, dy, dz)
    vorg = tuple(int(v) for v in rng.integers(-37, 23, size=3))
    fill = rng.uniform(0.05, 0.4)
    vox = np.where(rng.random((dims[2], dims[1], dims[0])) < fill, rng.integers(1, 14, size=(dims[2], dims[1], dims[0])), 0).astype(np.uint8)
    centre = np.array(vorg, dtype=np.float64) + np.array(dims) / 2.0
    n_lights = int(rng.integers(1, 5))
    lights = []
    for i in range(n_lights):
        pos = centre + rng.normal(size=3) * (np.array(dims) * (0.4 if i % 2 == 0 else 3.0))
        l = oracle.OrcLight()
        l.intensity = float(rng.uniform(1, 60))
        l.col[:] = [float(v) for v in rng.uniform(0.1, 1.2, size=3)]
        l.pos[:] = [float(np.float32(v)) for v in pos]
        lights.append(l)
    probe_count = tuple(int(v) for v in rng.integers(1, 4, size=3))
    side = int(rng.integers(2, 9))
    # field origin near the box, sometimes on exact lattice planes (fract == 0 at the first step)
    fo = centre + rng.normal(size=3) * np.array(dims) * 0.3
    if seed % 3 == 0:
        fo = np.round(fo)
    elif seed % 3 == 1:
        fo = np.round(fo * 2) / 2
    tile = (int(rng.integers(1, 7)), int(rng.integers(1, 7)))
    bounces = int(rng.integers(1, 9))
    screen = (48, 32)
    # every third seed also draws the optional modes: hysteresis blend, distance moments, Chebyshev
    # weight, a debug integrator, probe markers
    modes = dict(hysteresis=None, distance_scale=None, chebyshev=False, render_mode=0, visualize_probes=False)
    if seed % 3 == 2:
        modes = dict(hysteresis=float(rng.uniform(0.1, 0.95)), distance_scale=float(rng.choice([1.0, 4.0, side * 1.7])),
                     chebyshev=bool(rng.integers(0, 2)), render_mode=int(rng.integers(0, 6)), visualize_probes=bool(rng.integers(0, 2)))
    sc = oracle.Scene(probe_count=probe_count, side_length=side, field_origin=tuple(float(np.float32(v)) for v in fo), rx=tile[0], ry=tile[1],
                      lights=lights, scene=1, voxels=vox, vorg=vorg, max_bounces=bounces, screen=screen, **modes)
    rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, rng.normal(size=(tile[0] * tile[1], 3)).astype(np.float32)))
    cam_o = centre + rng.normal(size=3) * np.array(dims) * 0.8
    cam = ddgi_b200.Camera(screen[0] / float(screen[1]), tuple(float(v) for v in cam_o), tuple(float(v) for v in rng.uniform(-60, 60, size=3))).get_data()
    return dict(sc=sc, rays=rays, vox=vox, vorg=vorg, lights=lights, probe_count=probe_count, side=side, fo=fo, tile=tile, bounces=bounces,
                screen=screen, cam=cam, cam_o=cam_o, modes=modes)


SEEDS = list(range(1, 25))


@pytest.mark.parametrize("seed", SEEDS)
def test_engine_headers_on_random_scenes(seed):
    c = random_case(seed)
    sc, rays = c["sc"], c["rays"]
    W, H = sc.tex_size
    with np.errstate(all="ignore"):
        tex = np.zeros((H, W), dtype=np.uint32)
        for _ in range(2):   # two frames: the second blends into the first when the hysteresis mode is drawn
            want = oracle.probe_update(sc, rays, tex=tex)
        frame = oracle.render_frame(sc, c["cam"], want[0], tex_distances=want[1])
    hs = util.hostsim()
    for variant in (0, 1, 2):
        alb, f32, lk, dist = np.zeros_like(want[0]), np.zeros_like(want[2]), np.zeros_like(want[3]), np.zeros_like(want[1])
        for _ in range(2):
            hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, 0, sc.num_rays, variant, alb.ctypes.data, f32.ctypes.data, lk.ctypes.data,
                                dist.ctypes.data)
        util.assert_lookups(lk, want[3], variant, f"seed {seed}")
        assert np.array_equal(f32.view(np.uint32), want[2].view(np.uint32)), f"seed {seed} variant {variant}: fp32 texels"
        assert np.array_equal(alb, want[0]) and np.array_equal(dist, want[1])
    w, h = c["screen"]
    got, gf32, glk = np.zeros((h, w), dtype=np.uint32), np.zeros((h, w, 4), dtype=np.float32), np.zeros((h, w), dtype=np.uint32)
    hs.sim_render_frame(C.byref(sc.p), c["cam"].ctypes.data, want[0].ctypes.data, want[1].ctypes.data, got.ctypes.data, gf32.ctypes.data,
                        glk.ctypes.data)
    assert np.array_equal(glk, frame[2]), f"seed {seed}: pixel lookup counts"
    assert np.array_equal(gf32.view(np.uint32), frame[1].view(np.uint32)) and np.array_equal(got, frame[0])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS[::2])
def test_cuda_engine_on_random_scenes(seed):
    c = random_case(seed)
    sc, rays = c["sc"], c["rays"]
    W, H = sc.tex_size
    with np.errstate(all="ignore"):
        tex = np.zeros((H, W), dtype=np.uint32)
        for _ in range(2):
            want = oracle.probe_update(sc, rays, tex=tex)
        frame = oracle.render_frame(sc, c["cam"], want[0], tex_distances=want[1])
    w, h = c["screen"]
    m = c["modes"]
    with ddgi_b200.RVPT(w, h) as r:
        r.set_debug(True)
        r.render_settings.render_mode = m["render_mode"]
        r.render_settings.visualize_probes = 1 if m["visualize_probes"] else 0
        if m["hysteresis"] is not None:
            r.ir.hysteresis = m["hysteresis"]
        r.render_settings.scene = 1
        r.render_settings.max_bounces = c["bounces"]
        r.ir.probe_count[:] = c["probe_count"]
        r.ir.side_length = c["side"]
        r.ir.sqrt_rays_per_probe = c["tile"][0]
        r.ir.field_origin[:] = tuple(float(np.float32(v)) for v in c["fo"])
        r.ray_tile = c["tile"]
        r.lights = [capi.Light(l.intensity, tuple(l.col), tuple(l.pos)) for l in c["lights"]]
        r.upload_voxels(c["vox"], c["vorg"])
        r.set_probe_rays(rays)
        if m["hysteresis"] is not None:
            r.set_blend_mode(capi.BLEND_HYSTERESIS)
        if m["distance_scale"] is not None:
            r.set_distance_mode(capi.DISTANCE_MOMENTS, m["distance_scale"])
        if m["chebyshev"]:
            r.set_weight_mode(capi.WEIGHT_CHEBYSHEV)

        class FixedCamera:
            def get_data(self_inner):
                return c["cam"]

        r.scene_camera = FixedCamera()
        for variant in (0, 1, 2):
            r.set_kernel_variant(variant)
            r.update(advance_time=False)
            r.write_probe_texture(np.zeros((H, W), dtype=np.uint32), 0)
            r.probe_update()
            r.draw()
            r.sync()
            util.assert_lookups(r.read_lookup_counts(0), want[3], variant, f"seed {seed}")
            assert np.array_equal(r.read_probe_texture(0, capi.FMT_F32).view(np.uint32), want[2].view(np.uint32))
            assert np.array_equal(r.read_probe_texture(0), want[0]) and np.array_equal(r.read_probe_texture(1), want[1])
            assert np.array_equal(r.read_lookup_counts(1).reshape(h, w), frame[2])
            assert np.array_equal(r.read_frame(capi.FMT_F32).view(np.uint32), frame[1].view(np.uint32))
            assert np.array_equal(r.read_frame(), frame[0])
