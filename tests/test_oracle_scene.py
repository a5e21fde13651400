"""Oracle scene functions against facts read off the reference source, and the
stored-voxel extension against the literal procedural mode."""
import ctypes as C

import numpy as np

import util
import pytest

from oracle import oracle, ref

CFG = util.configs.CONFIGS


def _block(sc, x, y, z):
    c = np.array([x, y, z], dtype=np.float32)
    return oracle.load().orc_get_block_at(C.byref(sc.p), c.ctypes.data)


def test_cornell_walls_and_boxes():
    """assets/shaders/intersection.glsl:758-791."""
    sc = util.oracle_scene(CFG["cornell_2x2x2"], procedural=True)
    assert _block(sc, -10, 0, 15) == 2      # red wall x = -10
    assert _block(sc, 10, 0, 15) == 3       # green wall x = 10
    assert _block(sc, 0, 10, 15) == 5 and _block(sc, 0, -10, 15) == 5
    assert _block(sc, 0, 0, 25) == 5        # back wall
    assert _block(sc, 0, 0, 5) == 0         # front is open
    assert _block(sc, -3, -7, 13) == 5      # small box
    assert _block(sc, 4, -4, 16) == 5       # tall box
    assert _block(sc, 0, 0, 15) == 0
    assert _block(sc, -10, 10, 15) == 0     # |y| < 10 is strict on the side walls


def test_cave_structure():
    """assets/shaders/intersection.glsl:720-756."""
    sc = util.oracle_scene(CFG["cave_64"], procedural=True)
    assert _block(sc, 0, 18, 0) == 0            # above y = 17 is empty
    assert _block(sc, 0, 0, 0) == 0             # cavity
    assert _block(sc, 60, 0, 60) == 10          # rock outside the four spheres
    assert _block(sc, 0, -30, 0) in (11, 12)    # fbm floor band
    types = oracle.bake_scene(0, (64, 64, 64), (-32, -32, -32))
    assert set(np.unique(types)) <= set(range(14))
    assert {0, 10, 11}.issubset(set(np.unique(types)))
    assert any(t in np.unique(types) for t in (6, 7, 8, 9)), "mushrooms are inside the 64^3 crop"


def test_voxel_mode_equals_procedural_mode_on_cornell():
    """The whole Cornell geometry lies inside the 32^3 box, so the stored-voxel extension
    must reproduce the literal (procedural getBlockAt) reference path exactly."""
    cfg = CFG["cornell_3x3x3"]
    lit = util.oracle_scene(cfg, procedural=True)
    vox = util.oracle_scene(cfg)
    rays = oracle.generate_probe_rays(vox, oracle.generate_samples(8, 8))
    a = oracle.probe_update(lit, rays)
    b = oracle.probe_update(vox, rays)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    assert np.array_equal(a[3], b[3])
    cam = util.camera_block(cfg)
    small = util.small(cfg, screen=(128, 128))
    lit_s, vox_s = util.oracle_scene(small, procedural=True), util.oracle_scene(small)
    fa = oracle.render_frame(lit_s, cam, a[0])
    fb = oracle.render_frame(vox_s, cam, a[0])
    assert np.array_equal(fa[0], fb[0])


def test_generate_samples_matches_a_python_restatement():
    """src/rvpt/rvpt.cpp:1147-1173 with glibc rand(), seed 1, x jitter first."""
    libc = C.CDLL("libc.so.6")
    libc.rand.restype = C.c_int
    for s in (8, 16):
        got = oracle.generate_samples(s, s, reseed=True)
        libc.srand(1)
        f32 = np.float32
        inv = f32(1.0) / f32(s)
        i = 0
        for y in range(s):
            for x in range(s):
                j1 = f32(libc.rand()) / f32(2147483647)
                j2 = f32(libc.rand()) / f32(2147483647)
                su = f32(f32(x) + j1) * inv
                sv = f32(f32(y) + j2) * inv
                z = f32(1) - f32(f32(2) * su)
                assert got[i, 2] == z
                ang = f32(2.0 * 3.1415926 * float(sv))
                ring = np.sqrt(f32(1) - f32(z * z))
                assert abs(got[i, 0] - np.cos(ang) * ring) < 1e-6
                assert abs(got[i, 1] - np.sin(ang) * ring) < 1e-6
                i += 1
        n = np.linalg.norm(got, axis=1)
        assert np.abs(n - 1).max() < 1e-5
        # stratification: sample i lies in z-stratum x and phi-stratum y
        zs = ((1 - got[:, 2]) / 2 * s).astype(int).reshape(s, s)
        assert (zs == np.arange(s)[None, :]).all()


def test_probe_lattice_and_tile_offsets():
    """src/rvpt/rvpt.cpp:1190-1221: integer (dim-1)/2 centring, tile offset (i % s, i / s)."""
    cfg = CFG["cornell_2x2x2"]
    sc = util.oracle_scene(cfg)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8))
    assert rays.shape == (512, 12)
    # probe 0 of a 2x2x2 field: index - (2-1)/2 = index - 0 -> origin at field_origin
    assert tuple(rays[0, 0:3]) == (0.0, 0.0, 15.0)
    # probe 7 = (1,1,1) -> +side on every axis
    assert tuple(rays[7 * 64, 0:3]) == (15.0, 15.0, 30.0)
    assert tuple(rays[7 * 64 + 10, 8:11]) == (7.0, 2.0, 1.0)
    odd = util.oracle_scene(CFG["cornell_3x3x3"])
    r3 = oracle.generate_probe_rays(odd, oracle.generate_samples(8, 8))
    assert tuple(r3[13 * 64, 0:3]) == (0.0, 0.0, 15.0)  # centre probe of 3x3x3 sits on the origin


def test_probe_texture_invariants():
    cfg = CFG["cornell_3x3x3"]
    sc = util.oracle_scene(cfg)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8))
    alb, dist, f32, steps, oob = oracle.probe_update(sc, rays)
    assert (dist == 0).all()                       # probe_pass.comp:276,302
    assert ((alb >> 24) == 255).all()              # alpha = 1
    assert (f32[..., :3] >= 0).all()
    assert steps.max() <= 125 * 8 * 2              # B * (1 + L) marches of <= 125 steps
    assert steps.min() >= 1
    # rectangular tile reduces to the square case
    sq = util.oracle_scene(util.small(cfg, tile=(8, 8)))
    assert sq.tex_size == (72, 24)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("pc,side,s,org", [((3, 3, 3), 11, 8, (0.0, 0.0, 15.0)), ((2, 2, 2), 15, 8, (0.0, 0.0, 15.0)),
                                           ((9, 7, 9), 11, 20, (1.4, 0.0, 1.0)), ((2, 3, 4), 7, 5, (0.5, -2.0, 3.25))])
def test_host_ray_generator_against_the_reference_text_compiled_here(pc, side, s, org):
    """generate_samples + RVPT::generate_probe_rays (rvpt.cpp:1145-1224) and struct ProbeRay (probe.h) compiled
    verbatim against a glm stand-in (oracle/ref_glsl/build_ref.py).  The two rand() calls of rvpt.cpp:1161-1162
    are constructor arguments — their order is unspecified in C++; g++ runs the second first.  With that
    order the oracle's restatement is bit-identical to the compiled text; with PIN 5's order (x jitter
    first: what the fixtures and the engine use) only the jitter pairing, hence the directions, differ."""
    want = ref.generate_probe_rays(probe_count=pc, side_length=side, field_origin=org, s=s, reseed=True)
    sc = oracle.Scene(probe_count=pc, side_length=side, field_origin=org, rx=s, lights=[], scene=1, procedural=True)
    as_compiled = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True, y_first=True))
    assert np.array_equal(as_compiled.view(np.uint32), want.view(np.uint32))
    pinned = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
    assert np.array_equal(pinned[:, [0, 1, 2, 8, 9, 10]].view(np.uint32), want[:, [0, 1, 2, 8, 9, 10]].view(np.uint32))   # origins, probe / tile indices
    assert not np.array_equal(pinned[:, 4:7], want[:, 4:7])
    assert np.allclose(np.linalg.norm(pinned[:, 4:7], axis=1), 1.0, atol=1e-6)
