"""Shared test helpers: oracle scenes for the named configs, the hostsim loader."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import ddgi_b200
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = ddgi_b200._pkg
configs = __import__("importlib").import_module(pkg.__name__ + ".configs")

_hostsim = None


def hostsim():
    """TEST-ONLY host build of the engine's per-ray headers (tests/hostsim)."""
    global _hostsim
    if _hostsim is None:
        d = os.path.join(ROOT, "tests", "hostsim")
        subprocess.check_call(["make", "-s", "-C", d], stdout=subprocess.DEVNULL)
        lib = C.CDLL(os.path.join(d, "libhostsim.so"))
        P = C.POINTER(oracle.OrcParams)
        vp = C.c_void_p
        lib.sim_probe_update.argtypes = [P, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, vp, vp]
        lib.sim_render_frame.argtypes = [P, vp, vp, vp, vp, vp, vp]
        lib.sim_probe_update_oct.argtypes = [P, vp, C.c_int, vp, vp, vp]
        lib.sim_bake_scene.argtypes = [C.c_int, vp, vp, vp]
        lib.sim_pin_sincos.argtypes = [vp, C.c_int, vp, vp]
        lib.sim_pin_acos.argtypes = [vp, C.c_int, vp]
        _hostsim = lib
    return _hostsim


def synthetic_voxels(dims, permille, seed):
    """numpy restatement of the engine's integer-only synthetic cave baker."""
    dx, dy, dz = dims
    mind = min(dims)
    z, y, x = np.meshgrid(np.arange(dz, dtype=np.int64), np.arange(dy, dtype=np.int64),
                          np.arange(dx, dtype=np.int64), indexing="ij")
    inside = np.zeros((dz, dy, dx), dtype=bool)
    for cx, cy, cz, r in ((0, 0, 0, 20), (-16, -8, 10, 20), (13, 1, -19, 18), (-20, -15, -15, 21)):
        ex = 64 * x - 32 * dx - mind * cx
        ey = 64 * y - 32 * dy - mind * cy
        ez = 64 * z - 32 * dz - mind * cz
        inside |= (ex * ex + ey * ey + ez * ez) < (mind * r) ** 2
    i = ((z * dy + y) * dx + x).astype(np.uint64)
    h = (i & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ np.uint32(seed)
    with np.errstate(over="ignore"):
        h = (h ^ np.uint32(61)) ^ (h >> np.uint32(16))
        h = h * np.uint32(9)
        h = h ^ (h >> np.uint32(4))
        h = h * np.uint32(0x27D4EB2D)
        h = h ^ (h >> np.uint32(15))
        h = h ^ ((i >> np.uint64(32)).astype(np.uint32) * np.uint32(0x9E3779B9))
        h = h ^ (h << np.uint32(13))
        h = h ^ (h >> np.uint32(17))
        h = h ^ (h << np.uint32(5))
    solid = (h % np.uint32(1000)) < np.uint32(permille)
    t = np.where(solid, 2 + (h >> np.uint32(10)) % np.uint32(6), 0).astype(np.uint8)
    return np.where(inside, t, np.uint8(10)).astype(np.uint8)


def oracle_lights(cfg, time=0.0):
    if cfg["lights"] == "default":
        return oracle.default_lights(cfg["scene"])
    s = float(cfg.get("light_scale", 1.0))
    out = []
    for l in oracle.cave_lights4(time):
        m = oracle.OrcLight()
        m.intensity = np.float32(l.intensity) * np.float32(s)
        for a in range(3):
            m.col[a] = l.col[a]
            m.pos[a] = np.float32(l.pos[a]) * np.float32(s)
        out.append(m)
    return out


def oracle_voxels(cfg):
    v = cfg["voxels"]
    if v[0] == "bake":
        return oracle.bake_scene(cfg["scene"], v[1], v[2]), v[2]
    return synthetic_voxels(v[1], v[3], v[4]), v[2]


def oracle_scene(cfg, *, time=0.0, voxels=None, procedural=False, max_bounces=8) -> oracle.Scene:
    if voxels is None and not procedural:
        voxels, vorg = oracle_voxels(cfg)
    else:
        vorg = cfg["voxels"][2]
    rx, ry = cfg["tile"]
    return oracle.Scene(
        probe_count=cfg["probe_count"], side_length=cfg["side_length"], field_origin=cfg["field_origin"],
        rx=rx, ry=ry, lights=oracle_lights(cfg, time), scene=cfg["scene"], voxels=voxels, vorg=vorg,
        max_bounces=cfg.get("max_bounces", max_bounces), screen=cfg["screen"], procedural=procedural,
    )


def camera_block(cfg) -> np.ndarray:
    w, h = cfg["screen"]
    cam = ddgi_b200.Camera(w / float(h), cfg["camera"]["origin"], cfg["camera"]["rotation"])
    return cam.get_data()


def small(cfg, **over):
    c = dict(cfg)
    c.update(over)
    return c


def assert_lookups(got, want, variant, what=""):
    """Per-ray voxel-lookup counts against the reference algorithm's (oracle / reference shaders).
    Variants 0 and 1 perform exactly the reference's lookups (a different count = a different discrete
    path); variant 2 ends shadow-feeler marches behind their light (result-preserving early-out,
    csrc/ddgi_wavefront.cuh), so it may only ever perform fewer."""
    got, want = np.asarray(got), np.asarray(want)
    if variant == 2:
        assert (got <= want).all(), f"{what}: the early-out variant performed MORE lookups than the reference algorithm"
    else:
        assert np.array_equal(got, want), f"{what}: voxel lookup counts differ: a ray took another discrete path"
