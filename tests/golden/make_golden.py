#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE'S OWN SHADERS run on the CPU
(oracle/_ref/libddgi_ref.so = probe_pass.comp / compute_pass.comp transpiled where they lie
by oracle/ref_glsl/build_ref.py).  Needs /root/reference, so it runs in the build container
only; the fixtures are committed and travel to the GPU box.

    python oracle/ref_glsl/build_ref.py && python tests/golden/make_golden.py

Each fixture holds the inputs (ProbeRay list as the reference's generate_probe_rays lays it
out, camera block) and the reference outputs: probe texture (RGBA8 + the fp32 value handed to
imageStore), distance texture, per-invocation getBlockAt counts, frame (RGBA8 + fp32).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import oracle, ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (scene, probe_count, side_length, field_origin, s, screen, camera origin, camera rotation)
CASES = {
    # BASELINE configs[0] geometry (SURVEY.md 8d cfg 1) at a 128x128 frame
    "cornell_2x2x2": (1, (2, 2, 2), 15, (0.0, 0.0, 15.0), 8, (128, 128), (0.0, 0.0, -5.0), (0.0, 0.0, 0.0)),
    # odd-count twin, README.md:245-247 layout
    "cornell_3x3x3": (1, (3, 3, 3), 11, (0.0, 0.0, 15.0), 8, (128, 128), (0.0, 0.0, -5.0), (0.0, 0.0, 0.0)),
    # the reference's cave with its procedural textures, reference camera (rvpt.cpp:212-227)
    "cave_3x3x3": (0, (3, 3, 3), 7, (0.0, 0.0, 0.0), 8, (160, 96), (1.5, 2.0, -2.0), (-38.0, 36.0, 0.0)),
    # scene 2 (house), two lights
    "house_3x1x3": (2, (3, 1, 3), 9, (0.0, 0.0, 0.0), 6, (96, 64), (0.0, 0.0, -10.0), (0.0, 0.0, 0.0)),
}


# Fixtures for the pixel-pass modes the reference carries besides the DDGI frame, and for the
# lines it has commented out (restored by removing the comment markers, build_ref.py RESTORE):
# name -> (scene, probe_count, side_length, field_origin, s, screen, camera origin, camera rotation)
MODE_CASES = {
    "modes_cornell_3x3x3": (1, (3, 3, 3), 11, (0.0, 0.0, 15.0), 8, (64, 64), (0.0, 0.0, -5.0), (0.0, 0.0, 0.0)),
    "modes_cave_3x3x3": (0, (3, 3, 3), 7, (0.0, 0.0, 0.0), 6, (80, 48), (1.5, 2.0, -2.0), (-38.0, 36.0, 0.0)),
}
LIGHTS_TIME = 74.0  # render_settings.time of the animated-lights fixture (frame 37: time += 2 per frame)


def make_modes():
    import ddgi_b200

    for name, (scene, pc, side, org, s, screen, cam_o, cam_r) in MODE_CASES.items():
        kw = dict(scene=scene, probe_count=pc, side_length=side, field_origin=org, s=s)
        sc = oracle.Scene(probe_count=pc, side_length=side, field_origin=org, rx=s, lights=oracle.default_lights(scene),
                          scene=scene, procedural=True, literal_colors=True, screen=screen)
        rays = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
        alb, dist, _, _ = ref.probe_pass(rays=rays, **kw)
        cam = ddgi_b200.Camera(screen[0] / float(screen[1]), cam_o, cam_r).get_data()
        out = dict(scene=scene, probe_count=np.array(pc), side_length=side, field_origin=np.array(org, dtype=np.float32), s=s,
                   screen=np.array(screen), rays=rays, cam=cam, albedo=alb)
        # eval_integrator, compute_pass.comp:58-87: modes 1..5 and an out-of-range one (default branch = DDGI)
        for mode in (1, 2, 3, 4, 5, 9):
            f, f32, lk = ref.compute_pass(screen=screen, cam=cam, tex_albedo=alb, render_mode=mode, **kw)
            out[f"frame_mode{mode}"] = f
            out[f"frame_f32_mode{mode}"] = f32
            out[f"frame_lookups_mode{mode}"] = lk
        # probe markers (render_settings.visualize_probes) over the two integrators that draw them
        for mode in (0, 2):
            f, f32, _ = ref.compute_pass(screen=screen, cam=cam, tex_albedo=alb, render_mode=mode, visualize_probes=True, **kw)
            out[f"frame_markers_mode{mode}"] = f
            out[f"frame_f32_markers_mode{mode}"] = f32
        # `weight *= chebyshevWeight;` restored (intersection.glsl:1382): with the distance image the
        # reference writes (zeros) and with a random one
        rng = np.random.default_rng(20261017)
        dist_rand = rng.integers(0, 2 ** 32, size=alb.shape, dtype=np.uint32)
        out["distances_random"] = dist_rand
        for tag, d in (("zero", dist), ("random", dist_rand)):
            f, f32, _ = ref.compute_pass(screen=screen, cam=cam, tex_albedo=alb, tex_distances=d, chebyshev=True, **kw)
            out[f"frame_chebyshev_{tag}"] = f
            out[f"frame_f32_chebyshev_{tag}"] = f32
        # `update_lights();` restored in both main()s (probe_pass.comp:254, compute_pass.comp:174)
        la, _, lf32, llk = ref.probe_pass(rays=rays, animate_lights_time=LIGHTS_TIME, **kw)
        f, f32, _ = ref.compute_pass(screen=screen, cam=cam, tex_albedo=la, animate_lights_time=LIGHTS_TIME, **kw)
        out.update(lights_time=np.float32(LIGHTS_TIME), albedo_lights=la, albedo_f32_lights=lf32, lookups_lights=llk,
                   frame_lights=f, frame_f32_lights=f32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path)} bytes")


def main():
    import ddgi_b200

    if "--modes-only" in sys.argv:
        return make_modes()

    for name, (scene, pc, side, org, s, screen, cam_o, cam_r) in CASES.items():
        sc = oracle.Scene(probe_count=pc, side_length=side, field_origin=org, rx=s, lights=oracle.default_lights(scene),
                          scene=scene, procedural=True, literal_colors=True, screen=screen)
        rays = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
        alb, dist, f32, lk = ref.probe_pass(scene=scene, probe_count=pc, side_length=side, field_origin=org, s=s, rays=rays)
        cam = ddgi_b200.Camera(screen[0] / float(screen[1]), cam_o, cam_r).get_data()
        frame, frame_f32, frame_lk = ref.compute_pass(scene=scene, probe_count=pc, side_length=side, field_origin=org, s=s,
                                                      screen=screen, cam=cam, tex_albedo=alb, tex_distances=dist)
        # three frames of the reference's own (commented-out) hysteresis blend, probe_pass.comp:298-299 restored
        hyst = []
        prev = None
        for _ in range(3):
            prev = ref.probe_pass(scene=scene, probe_count=pc, side_length=side, field_origin=org, s=s, rays=rays,
                                  hysteresis=0.9, previous=prev)[0]
            hyst.append(prev)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, hysteresis=np.float32(0.9), albedo_hysteresis=np.stack(hyst), scene=scene, probe_count=np.array(pc), side_length=side, field_origin=np.array(org, dtype=np.float32),
                            s=s, screen=np.array(screen), rays=rays, cam=cam, albedo=alb, distances=dist, albedo_f32=f32,
                            lookups=lk, frame=frame, frame_f32=frame_f32, frame_lookups=frame_lk)
        print(f"{name}: {rays.shape[0]} rays, mean getBlockAt/ray {lk.mean():.1f}, frame {screen}, {os.path.getsize(path)} bytes")
    make_modes()


if __name__ == "__main__":
    main()
