#!/usr/bin/env python
"""Scheduling model of the probe-update kernel (no GPU): tests/hostsim re-enacts the warp loop of
probe_update_wavefront with the engine's own per-lane state functions on field_8 (the 1/64 twin of the
bench workload), 64 warps, rays taken in the cost-ordered schedule, and counts how often each piece of
code is issued and with how many lanes.  Costs per issue are warp instructions of the round-1 ncu split
(profiles/r1_probe_update_wavefront_e.txt: thread instructions per call of each state's code) so that
round 1's rule (split_hits = 1) lands near its measured 766 warp instructions per ray; the rules are then
compared in modelled warp instructions per ray.  A design tool, not a measurement.

    python profiles/policy_sim.py [workload]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import oracle  # noqa: E402

PIECES = ["MARCH", "SLOW", "LIGHT", "BOUNCE", "FEELER", "AIM", "SCATTER", "QUERY", "FETCH", "ROUND", "SWAP"]
# warp instructions per issue (round-1 capture e: BOUNCE_HIT 212 per call = light test + hit record + aim,
# FEELER_HIT 94 = light test + direct / ambient term, SCATTER 227, QUERY 83, FETCH 209, DDA step 68 + 12 of loop
# control, scheduler round 23; SWAP = a lane changing the ray it marches when it owns several)
COST = {"MARCH": 80.0, "SLOW": 170.0, "LIGHT": 30.0, "BOUNCE": 137.0, "FEELER": 64.0, "AIM": 45.0, "SCATTER": 227.0,
        "QUERY": 83.0, "FETCH": 209.0, "ROUND": 23.0, "SWAP": 30.0}


class PolicyOut(C.Structure):
    _fields_ = [("issues", C.c_uint64 * 11), ("lanes", C.c_uint64 * 11), ("makespan", C.c_double), ("busy", C.c_double)]


def bind(hs):
    hs.sim_wavefront_policy.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(PolicyOut)]
    return hs


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "field_8"
    cfg = util.configs.CONFIGS[name]
    sc = util.oracle_scene(cfg)
    rx, ry = cfg["tile"]
    rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True)))
    _, _, _, lk, _ = oracle.probe_update(sc, rays)
    n = rays.shape[0]
    slots = lk.reshape(-1, 32).max(axis=1)
    slot_order = np.argsort(-slots.astype(np.int64), kind="stable")
    order = np.ascontiguousarray((slot_order[:, None] * 32 + np.arange(32)[None, :]).reshape(-1).astype(np.uint32))
    hs = bind(util.hostsim())
    cost = np.array([COST[p] for p in PIECES], dtype=np.float64)

    def run(split, march_min, k, n_warps=64):
        out = PolicyOut()
        hs.sim_wavefront_policy(C.byref(sc.p), rays.ctypes.data, order.ctypes.data, n, n_warps, split, march_min, k, cost.ctypes.data, C.byref(out))
        return out

    print(f"{name}: {n} rays, mean lookups {lk.mean():.1f}")
    print(f"{'rule':46s} {'warp-inst/ray':>13s} {'vs r1':>6s}   lanes per issue: " + " ".join(f"{p[:6]:>6s}" for p in PIECES[:9]))
    ref = None
    for label, split, mm, k, nw in [("round 1: bounce / feeler hits separate, mm 16", 1, 16, 1, 64),
                                    ("merged HIT state, march_min 16", 0, 16, 1, 64), ("merged HIT state, march_min 12", 0, 12, 1, 64),
                                    ("merged HIT state, march_min 20", 0, 20, 1, 64), ("merged HIT state, march_min 24", 0, 24, 1, 64),
                                    ("merged, 2 rays per lane, march_min 16", 0, 16, 2, 32), ("merged, 2 rays per lane, march_min 24", 0, 24, 2, 32),
                                    ("merged, 2 rays per lane, march_min 28", 0, 28, 2, 32),
                                    ("merged, 3 rays per lane, march_min 28", 0, 28, 3, 21), ("merged, 4 rays per lane, march_min 28", 0, 28, 4, 16)]:
        o = run(split, mm, k, nw)
        total = o.busy / n
        ref = ref or total
        lanes = " ".join(f"{(o.lanes[i] / o.issues[i]) if o.issues[i] else 0:6.1f}" for i in range(9))
        print(f"{label:46s} {total:13.1f} {total / ref:6.3f}   {lanes}")


if __name__ == "__main__":
    main()
