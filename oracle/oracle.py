"""ctypes loader of the CPU oracle (oracle/ddgi_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddgi_oracle.so")


class OrcLight(C.Structure):
    _fields_ = [("intensity", C.c_float), ("col", C.c_float * 3), ("pos", C.c_float * 3)]


class OrcParams(C.Structure):
    _fields_ = [
        ("scene_mode", C.c_int32),
        ("scene", C.c_int32),
        ("color_mode", C.c_int32),
        ("n_lights", C.c_int32),
        ("lights", OrcLight * 8),
        ("vdim", C.c_int32 * 3),
        ("vorg", C.c_int32 * 3),
        ("vox", C.c_void_p),
        ("palette", C.c_void_p),
        ("probe_count", C.c_int32 * 3),
        ("side_length", C.c_int32),
        ("rx", C.c_int32),
        ("ry", C.c_int32),
        ("field_origin", C.c_float * 3),
        ("max_bounces", C.c_int32),
        ("screen_width", C.c_int32),
        ("screen_height", C.c_int32),
        ("blend_mode", C.c_int32),
        ("hysteresis", C.c_float),
        ("render_mode", C.c_int32),
        ("visualize_probes", C.c_int32),
        ("weight_mode", C.c_int32),
        ("distance_mode", C.c_int32),
        ("distance_scale", C.c_float),
        ("layout", C.c_int32),
        ("oct", C.c_int32),
    ]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "ddgi_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER(OrcParams)
    vp = C.c_void_p
    lib.orc_sizeof_params.restype = C.c_int
    assert lib.orc_sizeof_params() == C.sizeof(OrcParams), "OrcParams layout mismatch"
    lib.orc_num_threads.restype = C.c_int
    lib.orc_default_lights.restype = C.c_int
    lib.orc_default_lights.argtypes = [C.c_int, C.POINTER(OrcLight)]
    lib.orc_cave_lights4.restype = C.c_int
    lib.orc_cave_lights4.argtypes = [C.POINTER(OrcLight)]
    lib.orc_update_lights_cave.argtypes = [C.POINTER(OrcLight), C.c_int, C.c_float, C.POINTER(OrcLight)]
    lib.orc_default_palette.argtypes = [vp]
    lib.orc_pin_sincos.argtypes = [vp, C.c_int, vp, vp]
    lib.orc_pin_acos.argtypes = [vp, C.c_int, vp]
    lib.orc_wang_hash.restype = C.c_uint32
    lib.orc_wang_hash.argtypes = [C.c_uint32]
    lib.orc_rand_sequence.argtypes = [C.c_uint32, C.c_int, vp]
    lib.orc_hemisphere.argtypes = [vp, C.c_uint32, vp]
    lib.orc_generate_samples.argtypes = [C.c_int, C.c_int, C.c_int, vp]
    lib.orc_generate_samples_order.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp]
    lib.orc_generate_probe_rays.argtypes = [P, vp, vp]
    lib.orc_probe_update.argtypes = [P, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, C.c_int]
    lib.orc_render_frame.argtypes = [P, vp, vp, vp, vp, vp, vp, C.c_int]
    lib.orc_probe_update_oct.argtypes = [P, vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_int]
    lib.orc_update_lights.argtypes = [C.c_int, C.POINTER(OrcLight), C.c_int, C.c_float, C.POINTER(OrcLight)]
    lib.orc_sample_probe.argtypes = [P, vp, C.c_int, vp, vp]
    lib.orc_get_block_at.restype = C.c_int
    lib.orc_get_block_at.argtypes = [P, vp]
    lib.orc_get_color_at.argtypes = [P, vp, C.c_int, vp, vp]
    lib.orc_bake_scene.argtypes = [C.c_int, vp, vp, vp, C.c_int]
    lib.orc_intersect_scene.argtypes = [P, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def default_palette() -> np.ndarray:
    pal = np.zeros(256 * 3, dtype=np.float32)
    load().orc_default_palette(_ptr(pal))
    return pal


def default_lights(scene: int):
    arr = (OrcLight * 8)()
    n = load().orc_default_lights(scene, arr)
    return [arr[i] for i in range(n)]


def cave_lights4(time: float):
    base = (OrcLight * 8)()
    load().orc_cave_lights4(base)
    out = (OrcLight * 8)()
    load().orc_update_lights_cave(base, 4, C.c_float(time), out)
    return [out[i] for i in range(4)]


def update_lights(scene: int, lights, time: float):
    """update_lights (probe_pass.comp:217-250 == compute_pass.comp:126-160) applied to a light table."""
    n = len(lights)
    base = (OrcLight * 8)(*lights)
    out = (OrcLight * 8)()
    load().orc_update_lights(scene, base, n, C.c_float(time), out)
    return [out[i] for i in range(n)]


class Scene:
    """Keeps the numpy arrays an OrcParams points into alive."""

    def __init__(self, *, probe_count, side_length, field_origin, rx, ry=None, lights, scene=1,
                 voxels=None, vorg=(0, 0, 0), palette=None, max_bounces=8, screen=(0, 0), procedural=False,
                 literal_colors=False, hysteresis=None, render_mode=0, visualize_probes=False, chebyshev=False,
                 distance_scale=None, oct=None):
        self.p = OrcParams()
        p = self.p
        p.scene_mode = 0 if procedural else 1
        p.scene = scene
        p.color_mode = 0 if literal_colors else 1
        p.n_lights = len(lights)
        for i, l in enumerate(lights):
            p.lights[i].intensity = l.intensity
            for a in range(3):
                p.lights[i].col[a] = l.col[a]
                p.lights[i].pos[a] = l.pos[a]
        self.vox = None
        if voxels is not None:
            self.vox = np.ascontiguousarray(voxels, dtype=np.uint8)  # [z, y, x]
            dz, dy, dx = self.vox.shape
            p.vdim[:] = (dx, dy, dz)
            p.vorg[:] = tuple(vorg)
            p.vox = self.vox.ctypes.data
        self.palette = np.ascontiguousarray(palette if palette is not None else default_palette(), dtype=np.float32)
        p.palette = self.palette.ctypes.data
        p.probe_count[:] = tuple(probe_count)
        p.side_length = side_length
        p.rx = rx
        p.ry = ry if ry is not None else rx
        p.field_origin[:] = tuple(field_origin)
        p.max_bounces = max_bounces
        p.screen_width, p.screen_height = screen
        p.blend_mode = 0 if hysteresis is None else 1   # None = the reference as shipped (blend commented out)
        p.hysteresis = 0.0 if hysteresis is None else float(hysteresis)
        p.render_mode = render_mode
        p.visualize_probes = 1 if visualize_probes else 0
        p.weight_mode = 1 if chebyshev else 0           # 1: `weight *= chebyshevWeight` restored (G:1382)
        p.distance_mode = 0 if distance_scale is None else 1  # None = the reference as shipped: distances = vec2(0)
        p.distance_scale = 1.0 if distance_scale is None else float(distance_scale)
        p.layout = 0 if oct is None else 1   # None = the reference's one-texel-per-ray tile
        p.oct = 0 if oct is None else int(oct)

    @property
    def num_rays(self):
        p = self.p
        return p.probe_count[0] * p.probe_count[1] * p.probe_count[2] * p.rx * p.ry

    @property
    def tex_size(self):
        p = self.p
        tw, th = (p.oct, p.oct) if p.layout == 1 else (p.rx, p.ry)
        return p.probe_count[0] * p.probe_count[2] * tw, p.probe_count[1] * th


def generate_samples(rx: int, ry: int, reseed: bool = True, y_first: bool = False) -> np.ndarray:
    """y_first=False: PIN 5 (x jitter drawn first), what every fixture and the engine use.  y_first=True: the order
    g++ gives the reference's own text (oracle/_ref: ref_generate_probe_rays)."""
    out = np.zeros((rx * ry, 3), dtype=np.float32)
    load().orc_generate_samples_order(rx, ry, 1 if reseed else 0, 1 if y_first else 0, _ptr(out))
    return out


def generate_probe_rays(sc: Scene, samples: np.ndarray) -> np.ndarray:
    rays = np.zeros((sc.num_rays, 12), dtype=np.float32)
    s = np.ascontiguousarray(samples, dtype=np.float32)
    load().orc_generate_probe_rays(C.byref(sc.p), _ptr(s), _ptr(rays))
    return rays


def probe_update(sc: Scene, rays: np.ndarray, k0: int = 0, k1: int | None = None, threads: int = 0, tex=None):
    """Returns (albedo RGBA8 [H,W], distance RGBA8 [H,W], fp32 [H,W,4], lookups [R], oob [R])."""
    W, H = sc.tex_size
    k1 = sc.num_rays if k1 is None else k1
    alb = np.zeros((H, W), dtype=np.uint32) if tex is None else tex
    dist = np.zeros((H, W), dtype=np.uint32)
    f32 = np.zeros((H, W, 4), dtype=np.float32)
    steps = np.zeros(sc.num_rays, dtype=np.uint32)
    oob = np.zeros(sc.num_rays, dtype=np.uint32)
    r = np.ascontiguousarray(rays, dtype=np.float32)
    load().orc_probe_update(C.byref(sc.p), _ptr(r), k0, k1, _ptr(alb), _ptr(dist), _ptr(f32), _ptr(steps), _ptr(oob), threads)
    return alb, dist, f32, steps, oob


def probe_update_oct(sc: Scene, rays: np.ndarray, tex=None, dist=None, p0: int = 0, p1: int | None = None, threads: int = 0):
    """Octahedral layout (no reference output exists for it).  Returns (albedo [H,W], distance [H,W], lookups [R])."""
    W, H = sc.tex_size
    n_probes = sc.p.probe_count[0] * sc.p.probe_count[1] * sc.p.probe_count[2]
    p1 = n_probes if p1 is None else p1
    alb = np.zeros((H, W), dtype=np.uint32) if tex is None else tex
    dst = np.zeros((H, W), dtype=np.uint32) if dist is None else dist
    steps = np.zeros(sc.num_rays, dtype=np.uint32)
    r = np.ascontiguousarray(rays, dtype=np.float32)
    load().orc_probe_update_oct(C.byref(sc.p), _ptr(r), p0, p1, _ptr(alb), _ptr(dst), _ptr(steps), threads)
    return alb, dst, steps


def fibonacci_samples(n: int) -> np.ndarray:
    """The spherical-Fibonacci set of ddgi_generate_fibonacci_rays, restated independently (fp64, rounded once)."""
    i = np.arange(n, dtype=np.float64)
    f = i * 0.61803398874989484820
    az = 6.283185307179586476925 * (f - np.floor(f))
    z = 1.0 - (2.0 * i + 1.0) / n
    ring = np.sqrt(1.0 - z * z)
    return np.stack([np.cos(az) * ring, np.sin(az) * ring, z], axis=1).astype(np.float32)


def render_frame(sc: Scene, cam: np.ndarray, tex_albedo: np.ndarray, threads: int = 0, tex_distances=None):
    """Returns (frame RGBA8 [h,w], fp32 [h,w,4], lookups [h,w])."""
    w, h = sc.p.screen_width, sc.p.screen_height
    frame = np.zeros((h, w), dtype=np.uint32)
    f32 = np.zeros((h, w, 4), dtype=np.float32)
    steps = np.zeros((h, w), dtype=np.uint32)
    c = np.ascontiguousarray(cam, dtype=np.float32)
    t = np.ascontiguousarray(tex_albedo, dtype=np.uint32)
    d = np.zeros_like(t) if tex_distances is None else np.ascontiguousarray(tex_distances, dtype=np.uint32)
    load().orc_render_frame(C.byref(sc.p), _ptr(c), _ptr(t), _ptr(d), _ptr(frame), _ptr(f32), _ptr(steps), threads)
    return frame, f32, steps


def bake_scene(scene: int, dims, origin, threads: int = 0) -> np.ndarray:
    out = np.zeros((dims[2], dims[1], dims[0]), dtype=np.uint8)
    d = np.asarray(dims, dtype=np.int32)
    o = np.asarray(origin, dtype=np.int32)
    load().orc_bake_scene(scene, _ptr(o), _ptr(d), _ptr(out), threads)
    return out
