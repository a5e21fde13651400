"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): Cornell 3x3x3, both kernel
variants, literal SSBO rays and generated rays, distance moments, hysteresis, an edit, the octahedral
layout with Fibonacci rays (TMA-staged blend), every render mode with probe markers, procedural colours,
two frames in flight.  No torch import (fast start).
    compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitize_small.py"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddgi_b200  # noqa: E402

capi = ddgi_b200.capi
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
cfg = dict(configs.CONFIGS["cornell_3x3x3"])
cfg["screen"] = (64, 48)
with ddgi_b200.RVPT(64, 48) as r:
    configs.apply(r, cfg)
    r.generate_probe_rays(reseed=True)
    rays = r.probe_rays
    r.set_debug(True)
    for variant in (0, 1, 2):
        r.set_kernel_variant(variant)
        for ssbo in (False, True):
            if ssbo:
                r.set_probe_rays(rays)
            else:
                r.generate_probe_rays(reseed=True)
            r.update(advance_time=False)
            r.draw()
        r.set_distance_mode(capi.DISTANCE_MOMENTS, 19.0)
        r.set_blend_mode(capi.BLEND_HYSTERESIS)
        r.set_weight_mode(capi.WEIGHT_CHEBYSHEV)
        r.render_settings.visualize_probes = 1
        for mode in range(6):
            r.render_settings.render_mode = mode
            r.update(advance_time=False)
            r.draw()
        r.render_settings.visualize_probes = 0
        r.render_settings.render_mode = 0
        r.edit_voxels(np.full((3, 5, 2), 4, dtype=np.uint8), (-2, -9, 9))
        r.set_double_buffer(True)
        host = np.zeros(r.probe_texture_size[::-1], dtype=np.uint32)
        for _ in range(3):
            r.update()
            r.probe_update()
            r.read_probe_texture_async(host.ctypes.data, host.nbytes, 0)
        r.read_wait()
        r.set_double_buffer(False)
        r.generate_fibonacci_rays()
        r.set_layout(capi.LAYOUT_OCTAHEDRAL, 6)
        r.update(advance_time=False)
        r.draw()
        r.draw()
        r.set_layout(capi.LAYOUT_RAY_TILE)
        r.set_blend_mode(capi.BLEND_OVERWRITE)
        r.set_weight_mode(capi.WEIGHT_LITERAL)
        r.set_distance_mode(capi.DISTANCE_ZERO, 1.0)
        r.generate_probe_rays(reseed=True)
    # the normal launches (no debug buffers: the instantiation without the lookup counter), the procedural colours,
    # and two frames in flight on the engine's own streams with asynchronous reads and a voxel edit in between
    r.set_debug(False)
    r.set_kernel_variant(2)
    for color in (capi.COLOR_LITERAL, capi.COLOR_PALETTE):
        r.set_color_mode(color)
        r.update(advance_time=False)
        r.draw()
    r.set_double_buffer(True)
    r.set_frames_in_flight(2)
    host2 = [np.zeros(r.probe_texture_size[::-1], dtype=np.uint32) for _ in range(4)]
    for f in range(4):
        r.update()
        if f == 2:
            r.edit_voxels(np.full((2, 2, 2), 5, dtype=np.uint8), (1, -8, 12))
        r.probe_update()
        r.read_probe_texture_async(host2[f].ctypes.data, host2[f].nbytes, 0)
        r.render_frame()
    r.read_wait()
    r.frame_fence()
    r.sync()
    r.set_frames_in_flight(1)
    r.set_double_buffer(False)
    r.set_debug(2)
    r.set_kernel_variant(1)
    r.update(advance_time=False)
    r.probe_update()
    r.sync()
    t = r.read_warp_times()
    print("ok:", r.launch_count, "kernel launches,", len(t), "warps timed, frame checksum", int(r.read_frame().astype(np.uint64).sum()))
