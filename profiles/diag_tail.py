#!/usr/bin/env python
"""Diagnostic (1 GPU): where the time of ONE rank's share of field_32 goes when the field is
dealt to `world` ranks — per-warp start / last-fetch / exit times of the persistent kernel
(debug level 2), for a few grid limits and with / without the cost-ordered schedule.
Not a bench value.   python profiles/diag_tail.py [workload] [world]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddgi_b200  # noqa: E402

configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
cfg = configs.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "field_32"]
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream()
r.stream = stream.cuda_stream
X, Y, Z = cfg["probe_count"]
n = X * Y * Z * cfg["tile"][0] * cfg["tile"][1]


def timed(reps=10):
    for _ in range(3):
        r.probe_update()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)  # cold L2, as bench.py times it
        a.record(stream)
        r.probe_update()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if len(sys.argv) > 3 and sys.argv[3] == "march_min":   # sweep of the scheduling knob on both shares
    for w in (1, world):
        r.set_probes_cyclic(0, w, 1)
        for mm in (8, 12, 14, 16, 18, 20, 24):
            r.set_tuning(mm)
            print(f"world {w} march_min {mm:2d}: {timed():7.3f} ms")
    sys.exit(0)
for w in (1, world):
    for sched, slot in ((1, 0), (1, 128), (1, 64), (1, 32), (0, 0)):
        for limit in (0,):
            r.set_auto_schedule(bool(sched))
            r.set_schedule_slot(slot)
            r.set_probes_cyclic(0, w, 1)
            r.set_grid_limit(limit)
            r.set_debug(False)
            ms = timed()
            r.set_debug(2)
            r.probe_update()
            torch.cuda.synchronize()
            t = r.read_warp_times().astype(np.int64)
            r.set_debug(False)
            t0 = t[:, 0].min()
            start, last, exit_ = (t[:, 0] - t0) / 1e6, (t[:, 1] - t0) / 1e6, (t[:, 2] - t0) / 1e6
            q = lambda a, p: float(np.percentile(a, p))
            print(f"world {w} sched {sched} slot {slot or 'probe'}: {ms:7.3f} ms ({n / w / ms / 1e3:7.1f} Mrays/s) warps {len(t)} | "
                  f"start p50 {q(start, 50):.3f} max {start.max():.3f} | last fetch p10 {q(last, 10):.3f} p50 {q(last, 50):.3f} max {last.max():.3f} | "
                  f"exit p10 {q(exit_, 10):.3f} p50 {q(exit_, 50):.3f} p90 {q(exit_, 90):.3f} max {exit_.max():.3f} | "
                  f"mean busy {float((exit_ - start).mean()):.3f}")
