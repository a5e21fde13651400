#!/usr/bin/env python
"""Diagnostic (1 GPU): how evenly probe ownership splits the WORK of a workload over `world` ranks - kernel time of every
rank's share (one GPU plays each rank in turn) and the share's voxel lookups, for round-robin ownership with blocks of
1 / 3 / 5 / 33 probes.  (field_32: 32 x 32 x 32 probes, p = y*1024 + z*32 + x: dealt one by one to 8 ranks, rank r holds
the probe planes x = r, r+8, r+16, r+24 - not a uniform sample of the field.)
    python profiles/diag_balance.py [workload=field_32] [world=8] [blocks=1,3,5,33]"""
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
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
blocks = [int(b) for b in (sys.argv[3] if len(sys.argv) > 3 else "1,3,5,33").split(",")]
cfg = workload_config(name)
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream()
r.stream = stream.cuda_stream
X, Y, Z = cfg["probe_count"]
n_probes = X * Y * Z
rpp = cfg["tile"][0] * cfg["tile"][1]
r.set_debug(True)
r.probe_update()
r.sync()
lk = r.read_lookup_counts(0).reshape(n_probes, rpp).sum(axis=1).astype(np.float64)
r.set_debug(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for block in blocks:
    ms, work = [], []
    for rank in range(world):
        r.set_probes_cyclic(rank, world, block)
        for _ in range(3):
            r.probe_update()
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1)
            a.record(stream)
            r.probe_update()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms.append(float(np.median(ts)))
        owner = (np.arange(n_probes) // block) % world
        work.append(lk[owner == rank].sum())
    ms, work = np.array(ms), np.array(work)
    print(f"{name} / {world} ranks, blocks of {block:3d} probes: kernel ms per rank " + " ".join(f"{v:.3f}" for v in ms) +
          f"  max / mean = {ms.max() / ms.mean():.3f};  lookups max / mean = {work.max() / work.mean():.3f}", flush=True)
r.close()
