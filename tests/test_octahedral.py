"""The textbook layout the north star names: spherical-Fibonacci ray sets and an octahedral
oct x oct tile per probe filled with a cosine-weighted, warp-shuffle-reduced mean of ALL the
probe's rays (DDGI_LAYOUT_OCTAHEDRAL).

PARITY UNPINNED for this mode: the reference ships the mapping (assets/shaders/octahedral.glsl)
but never includes it and has no code that fills or reads such a tile, so there is no reference
output.  The contract is the oracle's statement of the operation order (oracle/ddgi_oracle.c:
orc_probe_update_oct); the engine's headers (CPU tier) and the CUDA engine (GPU tier) must match
it bit for bit, and the mode must satisfy the properties the construction implies.
"""
import ctypes as C

import numpy as np
import pytest

import ddgi_b200
import util
from oracle import oracle

capi = ddgi_b200.capi
CFG = util.configs.CONFIGS


def scene(oct=8, name="cornell_3x3x3", screen=(64, 64), **kw):
    cfg = util.small(CFG[name], screen=screen)
    vox, vorg = util.oracle_voxels(cfg)
    rx, ry = cfg["tile"]
    sc = oracle.Scene(probe_count=cfg["probe_count"], side_length=cfg["side_length"], field_origin=cfg["field_origin"], rx=rx, ry=ry,
                      lights=util.oracle_lights(cfg), scene=cfg["scene"], voxels=vox, vorg=vorg, screen=screen, oct=oct,
                      distance_scale=kw.pop("distance_scale", 19.0), **kw)
    return cfg, sc


def unpack(tex):
    return np.stack([(tex >> s) & 255 for s in (0, 8, 16)], axis=-1).astype(np.float64) / 255.0


def test_fibonacci_set_is_a_uniform_unit_set():
    d = oracle.fibonacci_samples(256).astype(np.float64)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)
    assert np.abs(d.mean(axis=0)).max() < 2e-3            # balanced
    assert np.all(np.diff(d[:, 2]) < 0)                   # z strictly decreasing: one ray per z band
    # nearest-neighbour angles are all alike (no clumps): min / max within a factor of 2.5
    cosn = np.sort(d @ d.T, axis=1)[:, -2]
    ang = np.arccos(np.clip(cosn, -1, 1))
    assert ang.max() / ang.min() < 2.5


def test_oracle_octahedral_properties():
    cfg, sc = scene(oct=8)
    rays = oracle.generate_probe_rays(sc, oracle.fibonacci_samples(64))
    alb, dist, lk = oracle.probe_update_oct(sc, rays)
    W, H = sc.tex_size
    assert alb.shape == (H, W) == (3 * 8, 9 * 8)
    # the per-ray lookup counts are those of the one-texel-per-ray layout: the trace is the same
    sc0 = oracle.Scene(probe_count=cfg["probe_count"], side_length=cfg["side_length"], field_origin=cfg["field_origin"], rx=8,
                       lights=util.oracle_lights(cfg), scene=cfg["scene"], voxels=sc.vox, vorg=cfg["voxels"][2], distance_scale=19.0)
    ray_tex, ray_dist, _, lk0, _ = oracle.probe_update(sc0, rays)
    assert np.array_equal(lk, lk0)
    # a texel is a convex combination of the probe's ray radiances: inside their range (up to the 8-bit rounding)
    e, r = unpack(alb), unpack(ray_tex)
    for p in range(27):
        ty, tx = (p // 9) * 8, (p % 9) * 8
        te, tr = e[ty:ty + 8, tx:tx + 8].reshape(-1, 3), r[ty:ty + 8, tx:tx + 8].reshape(-1, 3)
        assert (te.min(axis=0) >= tr.min(axis=0) - 1 / 255).all() and (te.max(axis=0) <= tr.max(axis=0) + 1 / 255).all()
    # distance plane: d in (0, 1], and E[d^2] >= E[d]^2 (Jensen) up to quantisation
    d = unpack(dist)
    assert (d[..., 0] > 0).all() and (d[..., 0] <= 1).all()
    assert (d[..., 1] + 2 / 255 >= d[..., 0] ** 2).all()
    # hysteresis: three frames converge monotonically towards the overwrite result
    cfg, sch = scene(oct=8, hysteresis=0.5)
    a = np.zeros_like(alb)
    dd = np.zeros_like(alb)
    prev = None
    for _ in range(3):
        oracle.probe_update_oct(sch, rays, tex=a, dist=dd)
        err = np.abs(unpack(a) - e).max()
        assert prev is None or err <= prev + 1 / 255
        prev = err
    assert prev < 0.2


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("oct,n", [(8, 64), (5, 96)])
def test_engine_headers_octahedral_match_oracle(variant, oct, n):
    """tests/hostsim: the engine's trace + ddgi_octahedral.cuh with the warp's lanes emulated in order."""
    cfg, sc = scene(oct=oct)
    sc.p.rx, sc.p.ry = (8, 8) if n == 64 else (8, 12)
    rays = oracle.generate_probe_rays(sc, oracle.fibonacci_samples(n))
    want = oracle.probe_update_oct(sc, rays)
    hs = util.hostsim()
    alb, dist = np.zeros_like(want[0]), np.zeros_like(want[0])
    lk = np.zeros(sc.num_rays, dtype=np.uint32)
    rays = np.ascontiguousarray(rays)
    hs.sim_probe_update_oct(C.byref(sc.p), rays.ctypes.data, variant, alb.ctypes.data, dist.ctypes.data, lk.ctypes.data)
    assert np.array_equal(lk, want[2])
    assert np.array_equal(alb, want[0]) and np.array_equal(dist, want[1])
    # pixel pass over the octahedral tiles, with the Chebyshev weight reading the distance tiles
    sc.p.weight_mode = 1
    cam = util.camera_block(cfg)
    f = oracle.render_frame(sc, cam, want[0], tex_distances=want[1])
    w, h = sc.p.screen_width, sc.p.screen_height
    frame, f32 = np.zeros((h, w), dtype=np.uint32), np.zeros((h, w, 4), dtype=np.float32)
    hs.sim_render_frame(C.byref(sc.p), cam.ctypes.data, want[0].ctypes.data, want[1].ctypes.data, frame.ctypes.data, f32.ctypes.data, None)
    assert np.array_equal(f32.view(np.uint32), f[1].view(np.uint32)) and np.array_equal(frame, f[0])
    assert (frame != 0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name,oct,tile", [("cornell_3x3x3", 8, (8, 8)), ("cornell_3x3x3", 5, (8, 12)), ("field_8", 8, (16, 16))])
def test_cuda_octahedral_matches_oracle(name, oct, tile, variant):
    cfg, sc = scene(oct=oct, name=name, screen=(64, 64), hysteresis=0.7)
    sc.p.rx, sc.p.ry = tile
    n = tile[0] * tile[1]
    rays = oracle.generate_probe_rays(sc, oracle.fibonacci_samples(n))
    cam = util.camera_block(cfg)
    alb, dist = np.zeros(sc.tex_size[::-1], dtype=np.uint32), np.zeros(sc.tex_size[::-1], dtype=np.uint32)
    with ddgi_b200.RVPT(64, 64) as r:
        r.set_debug(True)
        util.configs.apply(r, cfg)
        r.ray_tile = tile
        r.ir.hysteresis = 0.7
        r.generate_fibonacci_rays()
        assert np.array_equal(r.ray_samples.view(np.uint32), oracle.fibonacci_samples(n).view(np.uint32))
        r.set_layout(capi.LAYOUT_OCTAHEDRAL, oct)
        r.set_blend_mode(capi.BLEND_HYSTERESIS)
        r.set_distance_mode(capi.DISTANCE_ZERO, 19.0)   # (the scale is what the octahedral distance plane uses)
        r.set_weight_mode(capi.WEIGHT_CHEBYSHEV)
        r.set_kernel_variant(variant)
        r.update(advance_time=False)
        assert r.probe_texture_size == sc.tex_size
        sc.p.weight_mode = 1
        for frame_no in range(2):   # two frames: the second blends into the first
            _, _, lk = oracle.probe_update_oct(sc, rays, tex=alb, dist=dist)
            r.draw()
            r.sync()
            util.assert_lookups(r.read_lookup_counts(0), lk, variant)
            assert np.array_equal(r.read_probe_texture(0), alb), f"albedo plane, frame {frame_no}"
            assert np.array_equal(r.read_probe_texture(1), dist), f"distance plane, frame {frame_no}"
            want = oracle.render_frame(sc, cam, alb, tex_distances=dist)
            assert np.array_equal(r.read_frame(capi.FMT_F32).view(np.uint32), want[1].view(np.uint32))
            assert np.array_equal(r.read_frame(), want[0])
        # back to the reference layout: the textures are re-created at the ray-tile size
        r.set_layout(capi.LAYOUT_RAY_TILE)
        assert r.probe_texture_size == (sc.p.probe_count[0] * sc.p.probe_count[2] * tile[0], sc.p.probe_count[1] * tile[1])
