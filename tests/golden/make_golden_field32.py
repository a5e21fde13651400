#!/usr/bin/env python
"""Generates tests/golden/field_32.json: BASELINE.json configs[3] (32^3 probes x 256 rays, 4 moving lights, 512^3 synthetic
voxels, 1080p) at its FULL size from the CPU ORACLE (oracle/ddgi_oracle.c, OpenMP) at render_settings.time = 6: CRC-32 and
SHA-256 of the whole albedo texture (16384 x 512 RGBA8) and of the whole 1080p frame, the checksum of row checksums, and the
voxel-lookup totals.  The workload's voxel field is synthetic (no reference scene), so the reference's own shaders
(oracle/_ref) cannot run it; the oracle is pinned to them on the reference's scenes (tests/test_golden_reference.py).
~1 minute on 8 threads.

    python tests/golden/make_golden_field32.py
"""
import hashlib
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TIME = 6.0


def row_checksum(tex):
    rows = tex.astype(np.uint64).sum(axis=1)
    return int((rows * (np.arange(tex.shape[0], dtype=np.uint64) + 1)).sum() % (1 << 61))


def main():
    cfg = util.configs.CONFIGS["field_32"]
    sc = util.oracle_scene(cfg, time=TIME)
    rx, ry = cfg["tile"]
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))
    t0 = time.time()
    alb, dist, _, lk, _ = oracle.probe_update(sc, rays)
    t1 = time.time()
    print(f"probe update: {rays.shape[0]} rays in {t1 - t0:.1f} s, {lk.mean():.4f} lookups per ray")
    frame, _, flk = oracle.render_frame(sc, util.camera_block(cfg), alb)
    print(f"frame: {time.time() - t1:.1f} s")
    out = {
        "workload": "field_32", "time": TIME, "probe_rays": int(rays.shape[0]),
        "albedo_crc32": zlib.crc32(alb.tobytes()), "albedo_sha256": hashlib.sha256(alb.tobytes()).hexdigest(),
        "albedo_row_checksum": row_checksum(alb), "distance_all_zero": bool((dist == 0).all()),
        "lookups_sum": int(lk.sum(dtype=np.uint64)), "lookups_mean": float(lk.mean()),
        "frame_crc32": zlib.crc32(frame.tobytes()), "frame_sha256": hashlib.sha256(frame.tobytes()).hexdigest(),
        "frame_lookups_sum": int(flk.sum(dtype=np.uint64)),
    }
    with open(os.path.join(HERE, "field_32.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
