#!/usr/bin/env python
"""Runs N probe updates of rank 0's 1/world share of a workload (for ncu): python profiles/diag_share_run.py field_32 8 [n=6] [unit=1]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddgi_b200  # noqa: E402
from bench_support import workload_config  # noqa: E402

configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
cfg = workload_config(sys.argv[1] if len(sys.argv) > 1 else "field_32")
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
unit = int(sys.argv[4]) if len(sys.argv) > 4 else 1
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
r.stream = torch.cuda.current_stream().cuda_stream
r.set_probes_cyclic(0, world, unit)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(n):
    flush.fill_(1)
    r.probe_update()
    torch.cuda.synchronize()
r.close()
