#!/usr/bin/env python
"""Generates tests/golden/full_cave_128.npz: BASELINE.json configs[2] (cave, 16^3 probes x 256 rays, 1080p) at its FULL
size from the REFERENCE'S OWN SHADERS run on the CPU (oracle/_ref/libddgi_ref.so, see make_golden.py), with the
reference's procedural textures (the engine's DDGI_COLOR_LITERAL mode).  The full outputs (4 MiB texture, 8 MiB frame)
do not belong in the repository: the fixture keeps CRC-32 / SHA-256 of the whole albedo texture and of the whole frame,
the per-ray getBlockAt counts' sum, and for direct comparison the tiles of 96 sampled probes (RGBA8 + fp32 + lookups) and
a band of 64 frame rows.  Needs /root/reference (build container only); ~2 minutes, single thread.

    python oracle/ref_glsl/build_ref.py && python tests/golden/make_golden_cave128.py
"""
import hashlib
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import oracle, ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BAND = (512, 576)   # frame rows kept verbatim


def main():
    import ddgi_b200

    cfg = util.configs.CONFIGS["cave_128"]
    pc, side, org, s = cfg["probe_count"], cfg["side_length"], cfg["field_origin"], cfg["tile"][0]
    screen = cfg["screen"]
    sc = oracle.Scene(probe_count=pc, side_length=side, field_origin=org, rx=s, lights=oracle.default_lights(0), scene=0,
                      procedural=True, literal_colors=True, screen=screen)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
    kw = dict(scene=0, probe_count=pc, side_length=side, field_origin=org, s=s)
    t0 = time.time()
    alb, dist, f32, lk = ref.probe_pass(rays=rays, **kw)
    t1 = time.time()
    print(f"probe_pass.comp: {rays.shape[0]} rays in {t1 - t0:.1f} s, mean getBlockAt/ray {lk.mean():.2f}")
    assert (dist == 0).all()
    cam = ddgi_b200.Camera(screen[0] / float(screen[1]), cfg["camera"]["origin"], cfg["camera"]["rotation"]).get_data()
    frame, frame_f32, frame_lk = ref.compute_pass(screen=screen, cam=cam, tex_albedo=alb, tex_distances=dist, **kw)
    print(f"compute_pass.comp: {screen} in {time.time() - t1:.1f} s")
    X, Y, Z = pc
    n = s * s
    rng = np.random.default_rng(128)
    probes = np.array(sorted(set(rng.integers(0, X * Y * Z, size=93).tolist()) | {0, X * Y * Z - 1, (Y // 2 * Z + Z // 2) * X + X // 2}))
    tiles = alb.reshape(Y, s, X * Z, s).transpose(0, 2, 1, 3).reshape(X * Y * Z, s, s)
    tiles_f32 = f32.reshape(Y, s, X * Z, s, 4).transpose(0, 2, 1, 3, 4).reshape(X * Y * Z, s, s, 4)
    out = dict(
        probe_count=np.array(pc), side_length=side, field_origin=np.array(org, dtype=np.float32), s=s, screen=np.array(screen), cam=cam,
        albedo_crc32=np.uint32(zlib.crc32(alb.tobytes())), albedo_sha256=hashlib.sha256(alb.tobytes()).hexdigest(),
        frame_crc32=np.uint32(zlib.crc32(frame.tobytes())), frame_sha256=hashlib.sha256(frame.tobytes()).hexdigest(),
        lookups_sum=np.uint64(lk.sum(dtype=np.uint64)), frame_lookups_sum=np.uint64(frame_lk.sum(dtype=np.uint64)),
        probes=probes, tiles=tiles[probes], tiles_f32=tiles_f32[probes], tile_lookups=lk.reshape(X * Y * Z, n)[probes],
        band=np.array(BAND), frame_band=frame[BAND[0]:BAND[1]], frame_band_lookups=frame_lk[BAND[0]:BAND[1]].astype(np.uint16))
    path = os.path.join(HERE, "full_cave_128.npz")
    np.savez_compressed(path, **out)
    print(f"cave_128: {len(probes)} probe tiles + rows {BAND} of the frame, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
