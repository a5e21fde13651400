"""bench_support.py — the CPU legs of bench.py: `cpu_baseline` and `--impl reference`.

This is the ONE place outside tests/ and __graft_entry__.smoke() that executes oracle/:
it times the CPU restatement of the reference's path (the reference's GLSL cannot be
built or run: no glslang / Vulkan / lavapipe here or on the GPU box; DESIGN.md "Oracle")
on the box's host cores, on a bounded sample of the bench workload.  Nothing here is on
the product path.
"""
from __future__ import annotations

import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def workload_config(name: str) -> dict:
    import ddgi_b200

    configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
    if name.startswith("sweep_"):
        return configs.sweep_config(int(name.split("_")[1]))
    return configs.CONFIGS[name]


def _sample_rows(cfg, n_rows=None):
    """Probe rows the CPU legs trace: every (Y/4)-th row starting at 3/8 of that stride, i.e. spread over
    the whole field so that the sample's voxel lookups per ray are close to the field's mean (field_32:
    rows 3, 11, 19, 27 -> 104.9 lookups/ray against 106.4 for all 32 rows; the middle rows alone have 170)."""
    Y = cfg["probe_count"][1]
    n = max(1, min(Y, n_rows if n_rows else 4))
    stride = Y / float(n)
    return sorted({min(Y - 1, int(stride * i + 0.375 * stride)) for i in range(n)})


def _oracle_scene(cfg, voxels, time_value):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util

    return util.oracle_scene(cfg, time=time_value, voxels=voxels)


def host_threads():
    """All host cores, whatever OMP_NUM_THREADS says (torch.distributed.run exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _time_oracle(cfg, voxels, steps, warmup, time0=0.0, budget_s=200.0):
    """The oracle (OpenMP, `host_threads()` threads set explicitly) over the sample rows, `warmup` untimed +
    `steps` timed passes with the lights moving as in the GPU arm.  If the first pass says the whole run would
    exceed `budget_s`, the sample shrinks to 2 rows, then 1 (the steps stay the GPU arm's)."""
    from oracle import oracle

    rx, ry = cfg["tile"]
    X, Y, Z = cfg["probe_count"]
    sc = _oracle_scene(cfg, voxels, time0)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))
    per_row = X * Z * rx * ry
    threads = host_threads()
    W, H = sc.tex_size
    tex = np.zeros((H, W), dtype=np.uint32)
    rows = _sample_rows(cfg)

    def one_pass(sc, rows):
        t0 = time.perf_counter()
        total, n = 0.0, 0
        for y in rows:
            _, _, _, st, _ = oracle.probe_update(sc, rays, y * per_row, (y + 1) * per_row, threads=threads, tex=tex)
            total += float(st[y * per_row:(y + 1) * per_row].sum(dtype=np.float64))
            n += per_row
        return time.perf_counter() - t0, total / n

    times, traced, lookups = [], [], None
    for i in range(warmup + steps):
        sc = _oracle_scene(cfg, sc.vox, time0 + 2.0 * (i + 1))
        dt, lookups = one_pass(sc, rows)
        if i >= warmup:
            times.append(dt)
            traced.append(len(rows) * per_row)
        if i == 0 and dt * (warmup + steps) > budget_s and len(rows) > 1:
            rows = _sample_rows(cfg, 2 if dt * (warmup + steps) / 2 <= budget_s else 1)
    # (rays and seconds per step, averaged over the timed steps)
    return {"rays": float(np.mean(traced)), "seconds": float(np.mean(times)), "threads": threads, "rows": rows,
            "mean_lookups": lookups, "tex": tex}


def _reference_shaders_available(cfg):
    """oracle/_ref (the reference's own shaders transpiled to C++) has the reference's scenes, light
    tables and square ray tiles compiled in: it can run a workload only if that is what the workload is
    (the Cornell configs, and the cave with a square tile - cave_128, sweep_64 / 256 / 1024 - where the shaders run
    their own procedural, unbounded scene 0 with its procedural colours, intersection.glsl:699-756, 872-1047: the
    same algorithm and ray set as the workload the GPU arm times over the baked box; field_32 is a synthetic
    voxel field no shader text describes)."""
    from oracle import ref

    return ref.available() and cfg["scene"] in (0, 1) and cfg["lights"] == "default" and cfg["tile"][0] == cfg["tile"][1]


def _time_reference_shaders(cfg, steps, warmup, budget_s=150.0):
    """probe_pass.comp itself (transpiled, 1 thread: the shader's globals are process-wide) over the probe rays
    of the workload: all of them if `warmup + steps` passes fit in `budget_s` (the Cornell configs), else over every
    k-th probe (all rays of a sampled probe; k chosen from a first pass over every 64th probe), so that the run
    stays bounded whatever K / W the caller asks for."""
    from oracle import oracle, ref

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util

    s = cfg["tile"][0]
    sc = util.oracle_scene(cfg, procedural=True)
    rays_all = oracle.generate_probe_rays(sc, oracle.generate_samples(s, s, reseed=True))
    n_probes = rays_all.shape[0] // (s * s)

    def run(rays):
        t0 = time.perf_counter()
        out = ref.probe_pass(scene=cfg["scene"], probe_count=cfg["probe_count"], side_length=cfg["side_length"],
                             field_origin=cfg["field_origin"], s=s, rays=rays, max_bounces=cfg.get("max_bounces", 8))
        return time.perf_counter() - t0, out

    def every(k):
        return np.ascontiguousarray(rays_all.reshape(n_probes, s * s, -1)[k // 2::k].reshape(-1, rays_all.shape[1]))

    stride = 1
    if n_probes > 64:
        probe = every(64)
        dt, _ = run(probe)
        per_ray = dt / probe.shape[0]
        while stride < n_probes and per_ray * (rays_all.shape[0] / stride) * (warmup + steps) > budget_s:
            stride *= 2
    rays = rays_all if stride == 1 else every(stride)
    times = []
    for i in range(warmup + steps):
        dt, out = run(rays)
        if i >= warmup:
            times.append(dt)
    return {"rays": rays.shape[0], "seconds": float(np.mean(times)), "mean_lookups": float(out[3].mean()),
            "of": rays_all.shape[0], "stride": stride}


def _shader_sample(res):
    if res["stride"] == 1:
        return f"all {res['rays']} probe rays"
    return f"every {res['stride']}th probe = {res['rays']} of {res['of']} probe rays"


def cpu_baseline(rvpt, cfg, name):
    """The CPU side of the comparison on the host cores: the reference's own shaders (oracle/_ref,
    kind "reference") where they can run the workload, else the oracle ("port") over a bounded sample."""
    if _reference_shaders_available(cfg):
        res = _time_reference_shaders(cfg, steps=3, warmup=1, budget_s=30.0)
        return {"value": res["rays"] / res["seconds"], "unit": "probe-rays/s", "cores": 1, "kind": "reference",
                "sample": f"{_shader_sample(res)} per step, mean of 3 steps: the reference's probe_pass.comp transpiled "
                          f"to C++ (oracle/_ref), single thread",
                "mean_lookups_per_ray_in_sample": res["mean_lookups"]}
    X, Y, Z = cfg["probe_count"]
    vox = rvpt.read_voxels(cfg["voxels"][1])
    res = _time_oracle(cfg, vox, steps=3, warmup=1, budget_s=30.0)
    return {"value": res["rays"] / res["seconds"], "unit": "probe-rays/s", "cores": res["threads"], "kind": "port",
            "sample": f"probe rows {res['rows']} of {Y} (spread over the field) = {res['rays']} of "
                      f"{X*Y*Z*cfg['tile'][0]*cfg['tile'][1]} rays per step, mean of 3 steps, oracle/ddgi_oracle.c with OpenMP "
                      f"on {res['threads']} threads",
            "mean_lookups_per_ray_in_sample": res["mean_lookups"]}


def reference_arm(args, name):
    """`bench.py --impl reference`: same JSON shape, the oracle on the host cores."""
    cfg = workload_config(name)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util

    X, Y, Z = cfg["probe_count"]
    rx, ry = cfg["tile"]
    kind, cores = "port", None
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if _reference_shaders_available(cfg):
        res = _time_reference_shaders(cfg, steps=steps, warmup=warmup)
        kind, cores = "reference", 1
        sample = (f"{_shader_sample(res)} per step: the reference's own probe_pass.comp transpiled to C++ "
                  f"(oracle/_ref), single thread (its globals are process-wide)")
    else:
        vox, _ = util.oracle_voxels(cfg)
        res = _time_oracle(cfg, vox, steps=steps, warmup=warmup)
        cores = res["threads"]
        sample = (f"probe rows {res['rows']} of {Y} (spread over the field: {res['mean_lookups']:.1f} voxel lookups per ray in the "
                  f"sample) = {res['rays']} rays per step (the GLSL reference cannot be built: "
                  f"no glslang/Vulkan/lavapipe, and its transpiled shaders (oracle/_ref) have the reference's own scenes "
                  f"compiled in, not this workload's voxel field; this is the CPU oracle port, OpenMP, all host threads)")
    value = res["rays"] / res["seconds"]
    return {
        "impl": "reference", "metric": "probe_rays_per_s", "value": value, "unit": "probe-rays/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup,
        "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "probes": [X, Y, Z], "rays_per_probe": rx * ry, "probe_rays": X * Y * Z * rx * ry,
                   "voxels": list(cfg["voxels"][1]), "lights": 4 if cfg["lights"] == "cave4" else 1,
                   "max_bounces": cfg.get("max_bounces", 8), "resolution": list(cfg["screen"])},
        "cpu_baseline": {"value": value, "unit": "probe-rays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "probe-rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
