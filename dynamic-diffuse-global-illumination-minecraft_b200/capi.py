"""ctypes binding of libddgi_b200.so (include/ddgi.h).

This is the binding a Python caller of the C-ABI uses; the reference itself is C++
(src/rvpt/rvpt.cpp) and links the same symbols directly.  The library
is built in-tree by ``csrc/Makefile``; a missing library is an error — there is no
Python or CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DDGI_LIB overrides the library path (A/B builds while tuning); the default is the in-tree build
LIB_PATH = os.environ.get("DDGI_LIB") or os.path.join(_HERE, "libddgi_b200.so")

OK, E_INVALID, E_CUDA, E_STATE = 0, -1, -2, -3
FMT_RGBA8, FMT_F32 = 0, 1
COLOR_PALETTE, COLOR_LITERAL = 0, 1
BLEND_OVERWRITE, BLEND_HYSTERESIS = 0, 1
WEIGHT_LITERAL, WEIGHT_CHEBYSHEV = 0, 1
DISTANCE_ZERO, DISTANCE_MOMENTS = 0, 1
LAYOUT_RAY_TILE, LAYOUT_OCTAHEDRAL = 0, 1
RENDER_DDGI, RENDER_DIRECT, RENDER_INDIRECT, RENDER_COLOR, RENDER_NORMAL, RENDER_DEPTH = range(6)  # rvpt.h:25-31
MAX_LIGHTS = 8


class RenderSettings(C.Structure):
    """RVPT::RenderSettings, src/rvpt/rvpt.h:70-80 (32 bytes)."""

    _fields_ = [
        ("screen_width", C.c_int32),
        ("screen_height", C.c_int32),
        ("max_bounces", C.c_int32),
        ("camera_mode", C.c_int32),
        ("render_mode", C.c_int32),
        ("scene", C.c_int32),
        ("time", C.c_float),
        ("visualize_probes", C.c_int32),
    ]


class IrradianceField(C.Structure):
    """RVPT::IrradianceField, src/rvpt/rvpt.h:82-90 (48 bytes, std140)."""

    _fields_ = [
        ("probe_count", C.c_int32 * 3),
        ("side_length", C.c_int32),
        ("hysteresis", C.c_float),
        ("sqrt_rays_per_probe", C.c_int32),
        ("_pad0", C.c_int32 * 2),
        ("field_origin", C.c_float * 3),
        ("visualize", C.c_int32),
    ]


class ProbeRay(C.Structure):
    """struct ProbeRay, src/rvpt/probe.h:5-20 (48 bytes)."""

    _fields_ = [
        ("origin", C.c_float * 3),
        ("_p0", C.c_float),
        ("direction", C.c_float * 3),
        ("_p1", C.c_float),
        ("probe_info", C.c_float * 3),
        ("_p2", C.c_float),
    ]


class Light(C.Structure):
    """struct Light, assets/shaders/structs.glsl:54-59."""

    _fields_ = [("intensity", C.c_float), ("col", C.c_float * 3), ("pos", C.c_float * 3)]


assert C.sizeof(RenderSettings) == 32
assert C.sizeof(IrradianceField) == 48
assert C.sizeof(ProbeRay) == 48
assert C.sizeof(Light) == 28

_P = C.c_void_p
_I32 = C.c_int32
_SZ = C.c_size_t

# name -> (restype, argtypes); mirrors include/ddgi.h one to one
PROTOTYPES = {
    "ddgi_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "ddgi_destroy": (None, [_P]),
    "ddgi_last_error": (C.c_char_p, [_P]),
    "ddgi_version": (C.c_char_p, []),
    "ddgi_set_render_settings": (C.c_int, [_P, C.POINTER(RenderSettings)]),
    "ddgi_set_irradiance_field": (C.c_int, [_P, C.POINTER(IrradianceField)]),
    "ddgi_set_ray_tile": (C.c_int, [_P, _I32, _I32]),
    "ddgi_set_camera": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "ddgi_set_lights": (C.c_int, [_P, _I32, C.POINTER(Light)]),
    "ddgi_default_lights": (C.c_int, [_I32, C.POINTER(Light), C.POINTER(_I32)]),
    "ddgi_update_lights": (C.c_int, [_I32, C.c_float, C.POINTER(Light), _I32, C.POINTER(Light)]),
    "ddgi_cave_lights4": (C.c_int, [C.c_float, C.POINTER(Light)]),
    "ddgi_upload_voxels": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32), _P, _P]),
    "ddgi_bake_scene": (C.c_int, [_P, _I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "ddgi_bake_synthetic": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32), _I32, C.c_uint32]),
    "ddgi_set_color_mode": (C.c_int, [_P, _I32]),
    "ddgi_set_blend_mode": (C.c_int, [_P, _I32]),
    "ddgi_set_weight_mode": (C.c_int, [_P, _I32]),
    "ddgi_set_distance_mode": (C.c_int, [_P, _I32, C.c_float]),
    "ddgi_edit_voxels": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32), _P, _P]),
    "ddgi_read_voxels": (C.c_int, [_P, _P, _SZ]),
    "ddgi_generate_probe_rays": (C.c_int, [_P, _I32]),
    "ddgi_generate_fibonacci_rays": (C.c_int, [_P]),
    "ddgi_set_layout": (C.c_int, [_P, _I32, _I32]),
    "ddgi_set_ray_samples": (C.c_int, [_P, _P, _SZ]),
    "ddgi_get_ray_samples": (C.c_int, [_P, _P, _SZ]),
    "ddgi_set_probe_rays": (C.c_int, [_P, _P, _SZ]),
    "ddgi_get_probe_rays": (C.c_int, [_P, _P, _SZ]),
    "ddgi_num_probe_rays": (_SZ, [_P]),
    "ddgi_set_probe_rows": (C.c_int, [_P, _I32, _I32]),
    "ddgi_set_probe_rows_cyclic": (C.c_int, [_P, _I32, _I32, _I32]),
    "ddgi_set_probes_cyclic": (C.c_int, [_P, _I32, _I32, _I32]),
    "ddgi_probe_texture_device_ptr": (C.c_int, [_P, _I32, C.POINTER(_P), C.POINTER(_SZ)]),
    "ddgi_export_texture_handle": (C.c_int, [_P, _P]),
    "ddgi_export_texture_handles": (C.c_int, [_P, _P, C.POINTER(_I32)]),
    "ddgi_comm_unique_id": (C.c_int, [_P]),
    "ddgi_comm_init": (C.c_int, [_P, _P, _I32, _I32]),
    "ddgi_comm_destroy": (C.c_int, [_P]),
    "ddgi_exchange_allgather": (C.c_int, [_P, _P]),
    "ddgi_save_voxels": (C.c_int, [_P, C.c_char_p]),
    "ddgi_load_voxels": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(_I32), C.POINTER(_I32)]),
    "ddgi_save_checkpoint": (C.c_int, [_P, C.c_char_p, C.c_float]),
    "ddgi_load_checkpoint": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_float)]),
    "ddgi_set_sample_order": (C.c_int, [_P, _I32]),
    "ddgi_open_peers": (C.c_int, [_P, _I32, _P, _I32]),
    "ddgi_close_peers": (C.c_int, [_P]),
    "ddgi_exchange_barrier": (C.c_int, [_P, _P]),
    "ddgi_exchange_status": (C.c_int, [_P]),
    "ddgi_set_frame_band": (C.c_int, [_P, _I32, _I32]),
    "ddgi_frame_band_rows": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32)]),
    "ddgi_probe_update": (C.c_int, [_P, _P]),
    "ddgi_render_frame": (C.c_int, [_P, _P]),
    "ddgi_sync": (C.c_int, [_P]),
    "ddgi_probe_texture_size": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32)]),
    "ddgi_read_probe_texture": (C.c_int, [_P, _I32, _I32, _P, _SZ]),
    "ddgi_set_double_buffer": (C.c_int, [_P, _I32]),
    "ddgi_set_frames_in_flight": (C.c_int, [_P, _I32]),
    "ddgi_frame_fence": (C.c_int, [_P, _P]),
    "ddgi_last_update_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "ddgi_read_probe_texture_async": (C.c_int, [_P, _I32, _P, _SZ]),
    "ddgi_read_probe_texture_rows_async": (C.c_int, [_P, _I32, _I32, _I32, _P, _SZ]),
    "ddgi_read_wait": (C.c_int, [_P]),
    "ddgi_write_probe_texture": (C.c_int, [_P, _I32, _P, _SZ]),
    "ddgi_read_frame": (C.c_int, [_P, _I32, _P, _SZ]),
    "ddgi_set_debug": (C.c_int, [_P, _I32]),
    "ddgi_read_warp_times": (C.c_int, [_P, _P, _SZ, C.POINTER(_SZ)]),
    "ddgi_read_lookup_counts": (C.c_int, [_P, _I32, _P, _SZ]),
    "ddgi_set_kernel_variant": (C.c_int, [_P, _I32]),
    "ddgi_set_tuning": (C.c_int, [_P, _I32]),
    "ddgi_set_grid_limit": (C.c_int, [_P, _I32]),
    "ddgi_set_schedule_slot": (C.c_int, [_P, _I32]),
    "ddgi_set_auto_schedule": (C.c_int, [_P, _I32]),
    "ddgi_launch_count": (C.c_uint64, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libddgi_b200.so and declares every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built — run `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or __graft_entry__.build()); there is no fallback path"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
