"""GPU parity on BASELINE.json's own configurations at their FULL sizes (configs[2] cave_128, configs[4]
sweep_512 / sweep_1024 with 16x32 and 32x32 ray tiles, configs[3] field_32's pixel pass had no -m gpu test).

  * cave_128 with the reference's procedural textures, DIRECTLY against the reference's own shaders run on the CPU at
    full size (tests/golden/full_cave_128.npz, made by tests/golden/make_golden_cave128.py from oracle/_ref): CRC-32 / SHA-256 of
    the whole 4096 x 256 albedo texture and of the whole 1080p frame, 94 probe tiles (bytes, fp32 bits, getBlockAt counts)
    and a band of 64 frame rows compared value by value;
  * cave_128 as benched (flat palette), sweep_512, sweep_1024: the oracle on a sample of probes, on the engine's own voxels;
  * field_32: the whole 16384 x 512 texture and the whole 1080p frame against the oracle's CRC-32 / SHA-256 / row checksum
    (tests/golden/field_32.json, tests/golden/make_golden_field32.py).
"""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import ddgi_b200
import util
from ddgi_b200 import capi
from oracle import oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CFG = util.configs.CONFIGS


def tiles_of(tex, cfg):
    X, Y, Z = cfg["probe_count"]
    rx, ry = cfg["tile"]
    if tex.ndim == 2:
        return tex.reshape(Y, ry, X * Z, rx).transpose(0, 2, 1, 3).reshape(X * Y * Z, ry, rx)
    return tex.reshape(Y, ry, X * Z, rx, -1).transpose(0, 2, 1, 3, 4).reshape(X * Y * Z, ry, rx, -1)


def test_cave_128_full_size_against_the_reference_shaders():
    g = np.load(os.path.join(HERE, "golden", "full_cave_128.npz"))
    cfg = CFG["cave_128"]
    X, Y, Z = cfg["probe_count"]
    n = cfg["tile"][0] * cfg["tile"][1]
    with ddgi_b200.RVPT(*cfg["screen"]) as r:
        r.set_debug(True)
        util.configs.apply(r, cfg, bake=False)
        # the reference's scene is procedural and unbounded: the cave proper lies inside [-64,64)^3, but above y = 17
        # everything is empty and the rock's top face extends sideways without end, and rays of the upper probes
        # (up to y = 56) reach it as far out as they can march (125 cells): bake wide enough to hold every such hit
        r.bake_scene((384, 128, 384), (-192, -64, -192), scene=0)
        r.set_color_mode(capi.COLOR_LITERAL)
        r.generate_probe_rays(reseed=True)
        r.update(advance_time=False)
        assert np.array_equal(r.scene_camera.get_data().view(np.uint32), g["cam"].view(np.uint32))
        for variant in (1, 2):
            r.set_kernel_variant(variant)
            r.write_probe_texture(np.zeros(r.probe_texture_size[::-1], dtype=np.uint32))
            r.draw()
            r.sync()
            tex = r.read_probe_texture(0)
            lk = r.read_lookup_counts(0)
            t, tf = tiles_of(tex, cfg), tiles_of(r.read_probe_texture(0, capi.FMT_F32), cfg)
            probes = g["probes"]
            assert np.array_equal(t[probes], g["tiles"]), f"variant {variant}: sampled probe tiles"
            assert np.array_equal(tf[probes].view(np.uint32), g["tiles_f32"].view(np.uint32)), f"variant {variant}: fp32 values"
            assert zlib.crc32(tex.tobytes()) == int(g["albedo_crc32"]), f"variant {variant}: CRC of the whole texture"
            assert hashlib.sha256(tex.tobytes()).hexdigest() == str(g["albedo_sha256"])
            assert (r.read_probe_texture(1) == 0).all()
            if variant == 1:
                assert np.array_equal(lk.reshape(X * Y * Z, n)[probes], g["tile_lookups"])
                assert int(lk.sum(dtype=np.uint64)) == int(g["lookups_sum"])
            else:
                assert (lk.reshape(X * Y * Z, n)[probes] <= g["tile_lookups"]).all()
            frame = r.read_frame()
            b0, b1 = (int(v) for v in g["band"])
            assert np.array_equal(frame[b0:b1], g["frame_band"])
            assert zlib.crc32(frame.tobytes()) == int(g["frame_crc32"])
            assert hashlib.sha256(frame.tobytes()).hexdigest() == str(g["frame_sha256"])
            flk = r.read_lookup_counts(1).reshape(frame.shape)
            assert np.array_equal(flk[b0:b1].astype(np.uint16), g["frame_band_lookups"])
            assert int(flk.sum(dtype=np.uint64)) == int(g["frame_lookups_sum"])


@pytest.mark.parametrize("name", ["cave_128", "sweep_512", "sweep_1024"])
def test_probe_sample_parity_at_full_size(name):
    """configs[2] as benched (flat palette) and configs[4]'s rectangular 16x32 and 32x32 ray tiles: the oracle on 64
    probes drawn at random plus the field's corners and centre, every kernel variant; cave_128 also the 1080p frame."""
    cfg = CFG[name] if name in CFG else util.configs.sweep_config(int(name.split("_")[1]))
    X, Y, Z = cfg["probe_count"]
    rx, ry = cfg["tile"]
    n = rx * ry
    rng = np.random.default_rng(len(name) * 1000 + n)
    probes = sorted(set(rng.integers(0, X * Y * Z, size=64).tolist()) | {0, X * Y * Z - 1, (Y // 2 * Z + Z // 2) * X + X // 2})
    with ddgi_b200.RVPT(*cfg["screen"]) as r:
        r.set_debug(True)
        util.configs.apply(r, cfg)
        r.generate_probe_rays(reseed=True)
        r.update(advance_time=False)
        vox = r.read_voxels(cfg["voxels"][1])
        sc = util.oracle_scene(cfg, voxels=vox)
        rays = oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))
        W, H = sc.tex_size
        want = np.zeros((H, W), dtype=np.uint32)
        want_lk = {}
        for p in probes:
            _, _, _, steps, _ = oracle.probe_update(sc, rays, p * n, (p + 1) * n, tex=want)
            want_lk[p] = steps[p * n:(p + 1) * n].copy()
        want_tiles = tiles_of(want, cfg)
        full = None
        for variant in (0, 1, 2):
            r.set_kernel_variant(variant)
            r.write_probe_texture(np.zeros((H, W), dtype=np.uint32))
            r.probe_update()
            r.sync()
            tex = r.read_probe_texture(0)
            lk = r.read_lookup_counts(0).reshape(X * Y * Z, n)
            assert np.array_equal(tiles_of(tex, cfg)[probes], want_tiles[probes]), f"{name} variant {variant}"
            for p in probes:
                util.assert_lookups(lk[p], want_lk[p], variant, f"{name} probe {p}")
            assert (tex >> 24 == 255).all()
            full = tex if full is None else full
            assert np.array_equal(tex, full), "the variants agree on every texel"
        if name == "cave_128":
            # pixel pass at 1080p over the engine's (sample-verified) texture
            r.render_frame()
            r.sync()
            frame = r.read_frame()
            want_frame, _, want_flk = oracle.render_frame(sc, util.camera_block(cfg), full)
            hy = (cfg["screen"][1] // 16) * 16
            assert np.array_equal(frame[:hy], want_frame[:hy])
            assert (frame[hy:] == 0).all()
            assert np.array_equal(r.read_lookup_counts(1).reshape(frame.shape)[:hy], want_flk[:hy])


def test_field_32_whole_texture_and_frame_against_the_oracle():
    """configs[3], the metric's configuration, at t = 6: all 8 388 608 texels and the whole 1080p frame against the
    full-size oracle run recorded in tests/golden/field_32.json (no sampling)."""
    with open(os.path.join(HERE, "golden", "field_32.json")) as f:
        g = json.load(f)
    cfg = CFG["field_32"]
    with ddgi_b200.RVPT(*cfg["screen"]) as r:
        r.set_debug(True)
        util.configs.apply(r, cfg, time=g["time"])
        r.generate_probe_rays(reseed=True)
        r.update(advance_time=False)
        for variant in (1, 2):
            r.set_kernel_variant(variant)
            r.write_probe_texture(np.zeros(r.probe_texture_size[::-1], dtype=np.uint32))
            r.draw()
            r.sync()
            tex = r.read_probe_texture(0)
            assert zlib.crc32(tex.tobytes()) == g["albedo_crc32"], f"variant {variant}"
            assert hashlib.sha256(tex.tobytes()).hexdigest() == g["albedo_sha256"]
            rows = tex.astype(np.uint64).sum(axis=1)
            assert int((rows * (np.arange(tex.shape[0], dtype=np.uint64) + 1)).sum() % (1 << 61)) == g["albedo_row_checksum"]
            assert (r.read_probe_texture(1) == 0).all() and g["distance_all_zero"]
            lk = r.read_lookup_counts(0)
            if variant == 1:
                assert int(lk.sum(dtype=np.uint64)) == g["lookups_sum"]
            else:
                assert int(lk.sum(dtype=np.uint64)) <= g["lookups_sum"]
            frame = r.read_frame()
            assert zlib.crc32(frame.tobytes()) == g["frame_crc32"]
            assert hashlib.sha256(frame.tobytes()).hexdigest() == g["frame_sha256"]
            assert int(r.read_lookup_counts(1).sum(dtype=np.uint64)) == g["frame_lookups_sum"]
