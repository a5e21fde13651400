"""The scheduling-policy model behind profiles/r1_policy_model.md (tests/hostsim: sim_wavefront_policy,
sim_probe_update_pooled_stats) — a design tool, kept honest by two cheap checks: every rule processes every
ray exactly once (executions and lanes add up), and the pooled block logic ends with the right texture for
pool sizes other than the kernel's."""
import ctypes as C

import numpy as np

import util
from oracle import oracle


class PolicyOut(C.Structure):
    _fields_ = [("exec", C.c_uint64 * 8), ("lanes", C.c_uint64 * 8), ("passes", C.c_uint64), ("makespan", C.c_double), ("busy", C.c_double)]


def _scene():
    cfg = util.configs.CONFIGS["cornell_3x3x3"]
    sc = util.oracle_scene(cfg)
    rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8, reseed=True)))
    return sc, rays


def test_every_rule_resolves_every_query_once():
    sc, rays = _scene()
    n = rays.shape[0]
    want = oracle.probe_update(sc, rays)
    hs = util.hostsim()
    hs.sim_wavefront_policy.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(PolicyOut), C.c_int]
    order = np.arange(n, dtype=np.uint32)
    cost = np.ones(9, dtype=np.float64)
    lanes_ref = None
    for policy, mm, mo, group in ((0, 16, 0, 1), (0, 8, 0, 1), (1, 16, 0, 1), (2, 16, 8, 1), (3, 12, 20, 1), (4, 16, 4, 1), (0, 12, 0, 4)):
        out = PolicyOut()
        hs.sim_wavefront_policy(C.byref(sc.p), rays.ctypes.data, order.ctypes.data, n, 4, policy, mm, mo, cost.ctypes.data, C.byref(out), group)
        lanes = list(out.lanes)
        # lane-executions are a property of the rays, not of the rule: steps = lookups, one fetch per ray (+ the final empty ones)
        assert lanes[0] + lanes[6] == int(want[3].sum()), (policy, lanes)
        if lanes_ref is None:
            lanes_ref = lanes
        assert lanes[1:5] == lanes_ref[1:5], "queries / resolves / scatters per ray do not depend on the rule"
        assert out.busy > 0 and out.makespan * 4 >= out.busy


def test_pooled_block_logic_for_other_pool_sizes():
    sc, rays = _scene()
    want = oracle.probe_update(sc, rays)
    hs = util.hostsim()
    hs.sim_probe_update_pooled_stats.restype = C.c_uint64
    hs.sim_probe_update_pooled_stats.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_int, C.c_int]
    for slots, lockstep, keep in ((128, 0, 16), (128, 1, 24), (256, 1, 16), (64, 1, 8), (33, 1, 16)):
        alb, lk = np.zeros_like(want[0]), np.zeros_like(want[3])
        st = np.zeros(16, dtype=np.uint64)
        passes = hs.sim_probe_update_pooled_stats(C.byref(sc.p), rays.ctypes.data, rays.shape[0], 3, keep, alb.ctypes.data, lk.ctypes.data,
                                                  st.ctypes.data, slots, lockstep)
        assert passes > 0, "the pool did not drain"
        assert np.array_equal(alb, want[0]) and np.array_equal(lk, want[3]), (slots, lockstep, keep)
        assert int(st[11]) == int(want[3].sum())   # march lane-steps = voxel lookups
