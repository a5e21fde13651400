#!/usr/bin/env python
"""Diagnostic (1 GPU): kernel time of ONE rank's share of field_32 under probe-cyclic
ownership for world = 1, 2, 4, 8 — separates the kernel's fixed costs (launch, tail of the
persistent loop) from real multi-GPU effects.  Not a bench value."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddgi_b200  # noqa: E402

configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
cfg = configs.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "field_32"]
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream()
r.stream = stream.cuda_stream
X, Y, Z = cfg["probe_count"]
n = X * Y * Z * cfg["tile"][0] * cfg["tile"][1]
for world, sched in ((1, 0), (1, 1), (2, 1), (4, 1), (8, 0), (8, 1), (16, 1)):
    r.set_auto_schedule(bool(sched))
    for rank in sorted({0, world - 1}):
        r.set_probes_cyclic(rank, world, 1)
        for _ in range(3):
            r.probe_update()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            r.probe_update()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        med = ts[len(ts) // 2]
        print(f"sched {sched} world {world:2d} rank {rank:2d}: {med:7.3f} ms  rays {n // world:8d}  -> {n / world / med / 1e3:8.1f} Mrays/s")
