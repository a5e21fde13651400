#!/usr/bin/env python
"""Diagnostic (1 GPU): bench.py's end-to-end step on ONE rank's 1/world share (host buffers, H2D of the ray samples and uniforms,
asynchronous D2H of the rank's texture rows), with the wall time of every host call: where the e2e rate of a small share goes.
    python profiles/diag_e2e_share.py [workload=field_32] [world=8] [steps=60]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddgi_b200  # noqa: E402
from bench_support import workload_config  # noqa: E402

configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
name = sys.argv[1] if len(sys.argv) > 1 else "field_32"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
K = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cfg = workload_config(name)
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream()
r.stream = stream.cuda_stream
r.set_double_buffer(True)
r.set_probes_cyclic(0, world, 1)
X, Y, Z = cfg["probe_count"]
rx, ry = cfg["tile"]
n = X * Y * Z * rx * ry
W, H = r.probe_texture_size
rows = (0, H // world)
d2h = (rows[1] - rows[0]) * W * 4
lib = ddgi_b200.capi.load()
samples = torch.from_numpy(r.ray_samples).pin_memory()
host = [torch.empty(d2h, dtype=torch.uint8).pin_memory() for _ in range(2)]
frame = [0]
acc = {}


def t(name, f):
    t0 = time.perf_counter()
    v = f()
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return v


def e2e_step():
    frame[0] += 1
    r.render_settings.time = 2.0 * frame[0]
    r.lights = t("lights_for", lambda: configs.lights_for(cfg, r.render_settings.time))
    t("update", lambda: r.update(advance_time=False))
    t("set_ray_samples", lambda: lib.ddgi_set_ray_samples(r._ctx, samples.data_ptr(), rx * ry))
    t("probe_update", r.probe_update)
    t("read_async", lambda: r.read_probe_texture_rows_async(host[frame[0] & 1].data_ptr(), rows[0], rows[1], d2h, 0))


for fl in (1, 2):
    r.set_frames_in_flight(fl)
    for _ in range(5):
        e2e_step()
    r.read_wait()
    r.sync()
    acc.clear()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    t1 = time.perf_counter()
    r.read_wait()
    r.sync()
    t2 = time.perf_counter()
    print(f"{name} 1/{world} share, frames in flight {fl}: {(t2 - t0) / K * 1e3:.3f} ms per step end to end ({n / world * K / (t2 - t0) / 1e6:.0f} M probe-rays/s per rank), "
          f"host loop {(t1 - t0) / K * 1e3:.3f} ms per step: " + ", ".join(f"{k} {v / K * 1e3:.3f}" for k, v in acc.items()), flush=True)
r.set_frames_in_flight(1)
r.close()
