"""The workloads of BASELINE.json `configs`, as data (SURVEY.md §8d "Synthetic inputs").

Pure descriptions — no computation.  ``apply(rvpt, cfg)`` pushes one onto an RVPT host
object; tests build the same description for the oracle.
"""
from __future__ import annotations

import math

from . import capi


def _light(intensity, col, pos):
    return capi.Light(intensity, tuple(col), tuple(pos))


CONFIGS = {
    # configs[0]: Cornell box, 2x2x2 probes, 64 rays/probe, 512x512
    "cornell_2x2x2": dict(
        scene=1, voxels=("bake", (32, 32, 32), (-15, -15, 0)),
        probe_count=(2, 2, 2), side_length=15, field_origin=(0.0, 0.0, 15.0), tile=(8, 8),
        lights="default", screen=(512, 512), camera=dict(origin=(0.0, 0.0, -5.0), rotation=(0.0, 0.0, 0.0)),
    ),
    # odd-count twin (README.md:245-247 layout): the reference's cage indexing is only
    # self-consistent for odd probe counts
    "cornell_3x3x3": dict(
        scene=1, voxels=("bake", (32, 32, 32), (-15, -15, 0)),
        probe_count=(3, 3, 3), side_length=11, field_origin=(0.0, 0.0, 15.0), tile=(8, 8),
        lights="default", screen=(512, 512), camera=dict(origin=(0.0, 0.0, -5.0), rotation=(0.0, 0.0, 0.0)),
    ),
    # configs[1]: cave 64^3 voxels, 8^3 probes, 128 rays/probe (8x16 tile), 1080p
    "cave_64": dict(
        scene=0, voxels=("bake", (64, 64, 64), (-32, -32, -32)),
        probe_count=(8, 8, 8), side_length=7, field_origin=(0.0, 0.0, 0.0), tile=(8, 16),
        lights="default", screen=(1920, 1080), camera=dict(origin=(1.5, 2.0, -2.0), rotation=(-38.0, 36.0, 0.0)),
    ),
    # configs[2]: cave 128^3 voxels, 16^3 probes, 256 rays/probe, 1080p
    "cave_128": dict(
        scene=0, voxels=("bake", (128, 128, 128), (-64, -64, -64)),
        probe_count=(16, 16, 16), side_length=7, field_origin=(0.0, 0.0, 0.0), tile=(16, 16),
        lights="default", screen=(1920, 1080), camera=dict(origin=(1.5, 2.0, -2.0), rotation=(-38.0, 36.0, 0.0)),
    ),
    # configs[3]: 32^3 probes x 256 rays, 4 dynamic lights, 512^3 synthetic cave-like field
    "field_32": dict(
        scene=0, voxels=("synthetic", (512, 512, 512), (-256, -256, -256), 50, 0x9E3779B9),
        probe_count=(32, 32, 32), side_length=16, field_origin=(0.0, 0.0, 0.0), tile=(16, 16),
        lights="cave4", light_scale=8.0, screen=(1920, 1080),
        camera=dict(origin=(12.0, 16.0, -16.0), rotation=(-38.0, 36.0, 0.0)),
    ),
    # a 1/64-size twin of field_32 the CPU oracle finishes in seconds (parity tests)
    "field_8": dict(
        scene=0, voxels=("synthetic", (128, 128, 128), (-64, -64, -64), 50, 0x9E3779B9),
        probe_count=(8, 8, 8), side_length=16, field_origin=(0.0, 0.0, 0.0), tile=(16, 16),
        lights="cave4", light_scale=2.0, screen=(256, 256),
        camera=dict(origin=(3.0, 4.0, -4.0), rotation=(-38.0, 36.0, 0.0)),
    ),
}

# configs[4]: rays-per-probe sweep at 16^3 probes on the cave_128 scene
SWEEP_TILES = {64: (8, 8), 128: (8, 16), 256: (16, 16), 512: (16, 32), 1024: (32, 32)}


def sweep_config(rays_per_probe: int) -> dict:
    cfg = dict(CONFIGS["cave_128"])
    cfg["tile"] = SWEEP_TILES[rays_per_probe]
    return cfg


def lights_for(cfg: dict, time: float):
    """None = the reference table of cfg['scene']; 'cave4' = the 4-light cave table moved by
    update_lights (probe_pass.comp:219-235), scaled with the synthetic field."""
    import ctypes as C

    if cfg["lights"] == "default":
        return None
    lib = capi.load()
    arr = (capi.Light * 4)()
    rc = lib.ddgi_cave_lights4(C.c_float(time), arr)
    assert rc == 0
    s = float(cfg.get("light_scale", 1.0))
    out = []
    for l in arr:
        out.append(capi.Light(l.intensity * s, tuple(l.col), tuple(p * s for p in l.pos)))
    return out


def apply(rvpt, cfg: dict, *, bake: bool = True, time: float = 0.0):
    """Pushes a config onto an RVPT object (field, tile, screen, camera, lights, voxels)."""
    rs = rvpt.render_settings
    rs.screen_width, rs.screen_height = cfg["screen"]
    rs.scene = cfg["scene"]
    rs.max_bounces = cfg.get("max_bounces", 8)
    rs.time = time
    ir = rvpt.ir
    ir.probe_count[:] = cfg["probe_count"]
    ir.side_length = cfg["side_length"]
    ir.field_origin[:] = cfg["field_origin"]
    rx, ry = cfg["tile"]
    ir.sqrt_rays_per_probe = rx if rx == ry else int(math.isqrt(rx * ry))
    rvpt.ray_tile = (rx, ry)
    w, h = cfg["screen"]
    rvpt.scene_camera.aspect = w / float(h)
    rvpt.scene_camera.translation[:] = cfg["camera"]["origin"]
    rvpt.scene_camera.rotation[:] = cfg["camera"]["rotation"]
    rvpt.lights = lights_for(cfg, time)
    if bake:
        v = cfg["voxels"]
        if v[0] == "bake":
            rvpt.bake_scene(v[1], v[2], scene=cfg["scene"])
        else:
            rvpt.bake_synthetic(v[1], v[2], v[3], v[4])
