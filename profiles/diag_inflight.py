#!/usr/bin/env python
"""Diagnostic (1 GPU): ms per probe update of ONE rank's 1/world share of a workload with one and with two frames in
flight (ddgi_set_frames_in_flight) - what overlapping the drain of the persistent kernel with the next update's first
blocks is worth before any exchange cost.  K updates between two events, lights moved every step.  Not a bench value.

    python profiles/diag_inflight.py [workload=field_32] [worlds=1,2,4,8] [steps=40]
"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddgi_b200  # noqa: E402
from bench_support import workload_config  # noqa: E402

configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
name = sys.argv[1] if len(sys.argv) > 1 else "field_32"
worlds = [int(w) for w in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 40
limits = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "0").split(",")]  # resident blocks per SM (0 = all 7)
cfg = workload_config(name)
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream()
r.stream = stream.cuda_stream
r.set_double_buffer(True)
if os.environ.get('AUTO') == '0':
    r.set_auto_schedule(False)   # natural probe order instead of most expensive first
if os.environ.get('SLOT'):
    r.set_schedule_slot(int(os.environ['SLOT']))
X, Y, Z = cfg["probe_count"]
n = X * Y * Z * cfg["tile"][0] * cfg["tile"][1]
frame = [0]


mode = os.environ.get("FLUSH", "none")   # what runs on a side stream once per step (fl 2) / in order (fl 1)
side = torch.cuda.Stream()
buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def flush(fl):
    if mode == "none":
        return
    ctx = torch.cuda.stream(side) if fl == 2 else torch.cuda.stream(stream)
    with ctx:
        if mode == "fill256":
            buf.fill_(frame[0] & 255)
        elif mode == "fill160":
            buf[:160 << 20].fill_(frame[0] & 255)
        elif mode == "copy128":
            buf[:128 << 20].copy_(buf[128 << 20:])
        elif mode == "memset256":
            torch.cuda.cudart().cudaMemsetAsync(buf.data_ptr(), frame[0] & 255, 256 << 20, side.cuda_stream if fl == 2 else stream.cuda_stream)


def step():
    frame[0] += 1
    r.render_settings.time = 2.0 * frame[0]
    r.lights = configs.lights_for(cfg, r.render_settings.time)
    r.update(advance_time=False)
    r.probe_update()


for world, limit in [(w, l) for w in worlds for l in limits]:
    r.set_frames_in_flight(1)
    r.set_grid_limit(limit)
    r.set_probes_cyclic(0, world, 1)
    res, host = {}, {}
    for fl in (1, 2, 1, 2):
        r.set_frames_in_flight(fl)
        for _ in range(5):
            step()
        r.frame_fence()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        t0 = time.perf_counter()
        for _ in range(K):
            flush(fl)
            step()
        stream.wait_stream(side)
        host[fl] = (time.perf_counter() - t0) / K * 1e3
        r.frame_fence()
        b.record(stream)
        torch.cuda.synchronize()
        res.setdefault(fl, []).append(a.elapsed_time(b) / K)
    r.set_frames_in_flight(1)
    m1, m2 = min(res[1]), min(res[2])
    print(f"{os.path.basename(os.environ.get('DDGI_LIB', '') or 'default'):20s} host enqueue {host[1]:.3f} / {host[2]:.3f} ms per step; "
          f"auto {os.environ.get('AUTO', '1')} slot {os.environ.get('SLOT', '32')} flush {mode} grid limit {limit}: {name} 1/{world} share ({n // world} rays): one frame at a time {m1:.3f} ms, two in flight {m2:.3f} ms ({m1 / m2:.3f}x); "
          f"x{world} = {n / m1 / 1e3:.0f} -> {n / m2 / 1e3:.0f} M probe-rays/s if the exchange were free", flush=True)
r.close()
