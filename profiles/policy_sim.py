#!/usr/bin/env python
"""Scheduling-policy model of the probe-update kernel (no GPU): tests/hostsim re-enacts the warp
loop of probe_update_wavefront with the engine's own per-lane state functions on field_8 (the 1/64
twin of the bench workload), 64 warps, rays taken in the cost-ordered schedule, and counts how
often each state's code runs and with how many lanes.  Per-execution costs are calibrated so that
the kernel's own rule reproduces the ncu per-state instruction split (profiles/
r1_probe_update_wavefront_e.txt); other rules are then compared in modelled warp instructions per
ray and in makespan.  A design tool, not a measurement.

    python profiles/policy_sim.py
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

STATES = ["MARCH", "QUERY", "BOUNCE_HIT", "FEELER_HIT", "SCATTER", "FETCH", "MARCH_SLOW", "IDLE"]
# ncu capture e, warp instructions per ray by state (766 in all)
MEASURED = {"MARCH": 344.0, "QUERY": 70.5, "BOUNCE_HIT": 101.9, "FEELER_HIT": 42.1, "SCATTER": 98.8, "FETCH": 14.6, "MARCH_SLOW": 0.8,
            "scheduler": 94.2}
MARCH_LOOP = 14.0   # ballot + popc + compare + branch + the mode test, per march iteration (inside "scheduler")


class PolicyOut(C.Structure):
    _fields_ = [("exec", C.c_uint64 * 8), ("lanes", C.c_uint64 * 8), ("passes", C.c_uint64), ("makespan", C.c_double), ("busy", C.c_double)]


def main():
    cfg = util.configs.CONFIGS["field_8"]
    sc = util.oracle_scene(cfg)
    rx, ry = cfg["tile"]
    rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True)))
    _, _, _, lk, _ = oracle.probe_update(sc, rays)
    n = rays.shape[0]
    slots = lk.reshape(-1, 32).max(axis=1)
    slot_order = np.argsort(-slots.astype(np.int64), kind="stable")
    order = np.ascontiguousarray((slot_order[:, None] * 32 + np.arange(32)[None, :]).reshape(-1).astype(np.uint32))
    hs = util.hostsim()
    hs.sim_wavefront_policy.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(PolicyOut), C.c_int]
    n_warps = 64   # 2048 lanes: 64 rays per lane, the regime of field_32 on 4144 resident warps

    def run(policy, march_min=16, min_other=0, cost=None, group=1):
        out = PolicyOut()
        c = np.zeros(9, dtype=np.float64) if cost is None else np.ascontiguousarray(cost, dtype=np.float64)
        hs.sim_wavefront_policy(C.byref(sc.p), rays.ctypes.data, order.ctypes.data, n, max(1, n_warps // group), policy, march_min, min_other,
                                c.ctypes.data, C.byref(out), group)
        return out

    # calibration: the kernel's own rule with unit costs -> executions per ray -> cost per execution
    base = run(0, cost=np.ones(9))
    ex = {s: base.exec[i] / n for i, s in enumerate(STATES)}
    cost = np.zeros(9)
    for i, s in enumerate(STATES[:7]):
        cost[i] = MEASURED[s] / ex[s] if ex[s] > 0 else 0.0
    cost[0] += MARCH_LOOP
    passes = base.passes / n
    cost[8] = (MEASURED["scheduler"] - MARCH_LOOP * ex["MARCH"]) / passes
    print(f"field_8: {n} rays, mean lookups {lk.mean():.1f}; calibrated cost per execution: "
          + ", ".join(f"{s} {cost[i]:.0f}" for i, s in enumerate(STATES[:7])) + f", scheduler round {cost[8]:.0f}")
    print(f"{'rule':44s} {'warp-inst/ray':>13s} {'vs kernel':>9s} {'makespan':>9s}   lanes per execution: " + " ".join(f"{s[:6]:>6s}" for s in STATES[:6]))
    ref_total = None
    for name, policy, mm, mo in [("kernel: march_min 16", 0, 16, 0), ("march_min 12", 0, 12, 0), ("march_min 20", 0, 20, 0),
                                 ("fullest state, march included", 1, 16, 0),
                                 ("march_min 16, other states need >= 8 lanes", 2, 16, 8), ("march_min 16, other states need >= 12 lanes", 2, 16, 12),
                                 ("march_min 16, other states need >= 16 lanes", 2, 16, 16), ("march_min 12, other states need >= 12 lanes", 2, 12, 12),
                                 ("march_min 20, other states need >= 12 lanes", 2, 20, 12),
                                 ("hysteresis: start at 18, keep to 14", 3, 14, 18), ("hysteresis: start at 20, keep to 12", 3, 12, 20),
                                 ("hysteresis: start at 16, keep to 12", 3, 12, 16), ("hysteresis: start at 18, keep to 16", 3, 16, 18),
                                 ("drain states with >= 1 lane, then march_min 16", 4, 16, 1), ("drain states with >= 4 lanes", 4, 16, 4),
                                 ("drain states with >= 8 lanes", 4, 16, 8), ("drain >= 4, march_min 12", 4, 12, 4),
                                 ("drain >= 4, march_min 20", 4, 20, 4)]:
        o = run(policy, mm, mo, cost)
        total = o.busy / n
        ref_total = ref_total or total
        lanes = " ".join(f"{(o.lanes[i] / o.exec[i]) if o.exec[i] else 0:6.1f}" for i in range(6))
        print(f"{name:44s} {total:13.1f} {total / ref_total:9.3f} {o.makespan / (o.busy / n_warps):9.3f}   {lanes}")


    # upper bound on block-level compaction: the scheduling unit is a block of `group` warps whose rays
    # are regrouped by state for free, a state's code issued ceil(count / 32) times
    print()
    for move in (0.0, 60.0, 120.0):
        for group in (1, 2, 4, 8):
            if group == 1 and move > 0:
                continue
            best = None
            for mm in (16, 12, 8):
                c = cost.copy()
                c[7] = move   # instructions to move one warp-load of ray state pool <-> registers, per issue
                o = run(0, mm, 0, c, group)
                total = o.busy / n
                if best is None or total < best[0]:
                    best = (total, mm, o)
            total, mm, o = best
            lanes = " ".join(f"{(o.lanes[i] / o.exec[i]) if o.exec[i] else 0:6.1f}" for i in range(6))
            print(f"{'regrouping over %d warps, move cost %3.0f, march_min %d' % (group, move, mm):52s} {total:7.1f} {total / ref_total:7.3f}   {lanes}")


if __name__ == "__main__":
    main()
