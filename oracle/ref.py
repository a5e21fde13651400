"""ctypes loader of oracle/_ref/libddgi_ref.so — the reference's own compute shaders
(probe_pass.comp, compute_pass.comp + includes) transpiled to C++ by
oracle/ref_glsl/build_ref.py and run on the CPU.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py).  The reference's scene and
lights are compiled into its shaders: only scenes 0 (cave), 1 (Cornell), 2 (house) with the
reference's own light tables and procedural block / colour functions can run here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libddgi_ref.so")


class RefSettings(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("screen_width", "screen_height", "max_bounces", "camera_mode", "render_mode", "scene")] + [
        ("time", C.c_float), ("visualize_probes", C.c_int32)]


class RefField(C.Structure):
    _fields_ = [("probe_count", C.c_int32 * 3), ("side_length", C.c_int32), ("hysteresis", C.c_float),
                ("sqrt_rays_per_probe", C.c_int32), ("field_origin", C.c_float * 3)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        from . import oracle

        oracle.load()  # libddgi_ref.so resolves the pinned sin/cos/acos from the oracle library
        lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        lib.ref_probe_pass.argtypes = [C.POINTER(RefSettings), C.POINTER(RefField), vp, C.c_uint32, C.c_int, C.c_int, vp, vp, vp, vp]
        lib.ref_probe_pass_hysteresis.argtypes = lib.ref_probe_pass.argtypes
        lib.ref_probe_pass_lights.argtypes = lib.ref_probe_pass.argtypes
        lib.ref_compute_pass.argtypes = [C.POINTER(RefSettings), C.POINTER(RefField), vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
        lib.ref_compute_pass_lights.argtypes = lib.ref_compute_pass.argtypes
        if hasattr(lib, "ref_generate_probe_rays"):
            lib.ref_generate_probe_rays.restype = C.c_uint32
            lib.ref_generate_probe_rays.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, vp, C.c_uint32]
        lib.ref_compute_pass_chebyshev.argtypes = lib.ref_compute_pass.argtypes
        _lib = lib
    return _lib


def _params(scene, probe_count, side_length, field_origin, s, screen=(0, 0), max_bounces=8, time=0.0, render_mode=0,
            visualize_probes=False):
    rs = RefSettings(screen[0], screen[1], max_bounces, 0, render_mode, scene, time, 1 if visualize_probes else 0)
    f = RefField()
    f.probe_count[:] = tuple(probe_count)
    f.side_length = side_length
    f.hysteresis = 0.9
    f.sqrt_rays_per_probe = s
    f.field_origin[:] = tuple(field_origin)
    return rs, f


def generate_probe_rays(*, probe_count, side_length, field_origin, s, reseed=True) -> np.ndarray:
    """The reference's HOST ray generator — generate_samples + RVPT::generate_probe_rays, src/rvpt/rvpt.cpp:1145-1224,
    and struct ProbeRay, src/rvpt/probe.h — compiled from its own text against a glm stand-in.  float32 [R, 12]."""
    n = probe_count[0] * probe_count[1] * probe_count[2] * s * s
    out = np.zeros((n, 12), dtype=np.float32)
    got = load().ref_generate_probe_rays((C.c_int * 3)(*probe_count), side_length, s, (C.c_float * 3)(*field_origin), 1 if reseed else 0,
                                         out.ctypes.data, n)
    assert got == n
    return out


def probe_pass(*, scene, probe_count, side_length, field_origin, s, rays, max_bounces=8, hysteresis=None, previous=None,
               animate_lights_time=None):
    """Runs probe_pass.comp::main for every texel.  rays: float32 [R, 12] in the reference's
    ProbeRay layout.  Returns (albedo RGBA8 [H,W], distances RGBA8 [H,W], fp32 [H,W,4], lookups [R]).
    hysteresis=h runs the build with the reference's commented-out blend restored
    (probe_pass.comp:298-299) on top of the texture `previous` (zeros if None).
    animate_lights_time=t runs the build with `update_lights();` restored (probe_pass.comp:254) at
    render_settings.time = t."""
    assert hysteresis is None or animate_lights_time is None
    rs, f = _params(scene, probe_count, side_length, field_origin, s, max_bounces=max_bounces,
                    time=0.0 if animate_lights_time is None else float(animate_lights_time))
    if hysteresis is not None:
        f.hysteresis = float(hysteresis)
    W = probe_count[0] * probe_count[2] * s
    H = probe_count[1] * s
    r = np.ascontiguousarray(rays, dtype=np.float32)
    alb = np.zeros((H, W), dtype=np.uint32) if previous is None else np.array(previous, dtype=np.uint32, copy=True)
    dist = np.full((H, W), 0xdeadbeef, dtype=np.uint32)
    f32 = np.zeros((H, W, 4), dtype=np.float32)
    lk = np.zeros(r.shape[0], dtype=np.uint32)
    fn = load().ref_probe_pass if hysteresis is None else load().ref_probe_pass_hysteresis
    if animate_lights_time is not None:
        fn = load().ref_probe_pass_lights
    fn(C.byref(rs), C.byref(f), r.ctypes.data, r.shape[0], W, H, alb.ctypes.data, dist.ctypes.data, f32.ctypes.data, lk.ctypes.data)
    return alb, dist, f32, lk


def compute_pass(*, scene, probe_count, side_length, field_origin, s, screen, cam, tex_albedo, tex_distances=None, max_bounces=8,
                 render_mode=0, visualize_probes=False, chebyshev=False, animate_lights_time=None):
    """Runs compute_pass.comp::main for every dispatched pixel (render_mode selects the integrator,
    compute_pass.comp:58-87).  chebyshev=True runs the build with `weight *= chebyshevWeight;`
    restored (intersection.glsl:1382); animate_lights_time=t the build with `update_lights();`
    restored (compute_pass.comp:174) at render_settings.time = t.
    Returns (frame RGBA8 [h,w], fp32 [h,w,4], lookups [h,w])."""
    assert not (chebyshev and animate_lights_time is not None)
    rs, f = _params(scene, probe_count, side_length, field_origin, s, screen=screen, max_bounces=max_bounces,
                    time=0.0 if animate_lights_time is None else float(animate_lights_time), render_mode=render_mode,
                    visualize_probes=visualize_probes)
    w, h = screen
    t = np.ascontiguousarray(tex_albedo, dtype=np.uint32)
    H, W = t.shape
    d = np.zeros_like(t) if tex_distances is None else np.ascontiguousarray(tex_distances, dtype=np.uint32)
    c = np.ascontiguousarray(cam, dtype=np.float32)
    frame = np.zeros((h, w), dtype=np.uint32)
    f32 = np.zeros((h, w, 4), dtype=np.float32)
    lk = np.zeros((h, w), dtype=np.uint32)
    fn = load().ref_compute_pass
    if chebyshev:
        fn = load().ref_compute_pass_chebyshev
    if animate_lights_time is not None:
        fn = load().ref_compute_pass_lights
    fn(C.byref(rs), C.byref(f), c.ctypes.data, t.ctypes.data, d.ctypes.data, W, H, frame.ctypes.data,
                            f32.ctypes.data, lk.ctypes.data)
    return frame, f32, lk
