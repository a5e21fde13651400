#!/usr/bin/env python
"""Diagnostic (1 GPU): (a) histogram of the per-slot cost (max voxel lookups of a slot's 32 rays) and per-ray lookups of a
workload - how much of it is "uniform" (every march one step); (b) [solid] the same workload's field shape over an ALL-SOLID
voxel box: every ray uniform - the cost of a uniform ray on its own.   python profiles/diag_rock.py field_32 [solid] [n]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddgi_b200  # noqa: E402
from bench_support import workload_config  # noqa: E402

configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
name = sys.argv[1] if len(sys.argv) > 1 else "field_32"
solid = len(sys.argv) > 2 and sys.argv[2] == "solid"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = workload_config(name)
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
if solid:
    v = cfg["voxels"]
    dx, dy, dz = v[1]
    r.upload_voxels(np.full((dz, dy, dx), 1, dtype=np.uint8), v[2])
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
r.stream = torch.cuda.current_stream().cuda_stream
if not solid:
    r.set_debug(True)
    r.probe_update()
    r.sync()
    lk = r.read_lookup_counts(0)
    r.set_debug(False)
    slot = lk.reshape(-1, 32).max(axis=1)
    mb = cfg.get("max_bounces", 8)
    print(f"{name}: {lk.size} rays, mean lookups {lk.mean():.2f}; rays with exactly {2 * mb} lookups: {(lk == 2 * mb).mean():.4f}; "
          f"slots with max cost <= {2 * mb}: {(slot <= 2 * mb).mean():.4f}; slots <= {2 * mb + 8}: {(slot <= 2 * mb + 8).mean():.4f}; "
          f"lookups of the other rays: mean {lk[lk != 2 * mb].mean():.1f}")
    print("slot-cost percentiles:", np.percentile(slot, [1, 10, 25, 50, 60, 70, 80, 90, 99, 100]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(n):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r.probe_update()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"{name}{' all solid' if solid else ''}: kernel {np.median(ts[2:]):.3f} ms")
r.close()
