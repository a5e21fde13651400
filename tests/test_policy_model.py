"""The scheduling model behind profiles/policy_sim.py (tests/hostsim: sim_wavefront_policy) — a design
tool, kept honest by one cheap check: every rule traces every ray to the end exactly once (march steps =
voxel lookups, one resolve per query), whatever the rule, the threshold or the number of rays per lane."""
import ctypes as C

import numpy as np

import util
from oracle import oracle

PIECES = ["MARCH", "SLOW", "LIGHT", "BOUNCE", "FEELER", "AIM", "SCATTER", "QUERY", "FETCH", "ROUND", "SWAP"]


class PolicyOut(C.Structure):
    _fields_ = [("issues", C.c_uint64 * 11), ("lanes", C.c_uint64 * 11), ("makespan", C.c_double), ("busy", C.c_double)]


def test_every_rule_resolves_every_query_once():
    cfg = util.configs.CONFIGS["cornell_3x3x3"]
    sc = util.oracle_scene(cfg)
    rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8, reseed=True)))
    n = rays.shape[0]
    want = oracle.probe_update(sc, rays)
    hs = util.hostsim()
    hs.sim_wavefront_policy.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.POINTER(PolicyOut)]
    order = np.arange(n, dtype=np.uint32)
    cost = np.ones(11, dtype=np.float64)
    ref = None
    for split, mm, k in ((1, 16, 1), (0, 16, 1), (0, 8, 1), (1, 24, 1), (0, 16, 2), (1, 28, 3)):
        out = PolicyOut()
        hs.sim_wavefront_policy(C.byref(sc.p), rays.ctypes.data, order.ctypes.data, n, 4, split, mm, k, cost.ctypes.data, C.byref(out))
        lanes = dict(zip(PIECES, out.lanes))
        # lane-executions are a property of the rays, not of the rule
        assert lanes["MARCH"] + lanes["SLOW"] == int(want[3].sum()), (split, mm, k, lanes)
        assert lanes["LIGHT"] == lanes["BOUNCE"] + lanes["FEELER"]
        assert lanes["QUERY"] == lanes["LIGHT"], "one resolve per query"
        per_ray = (lanes["BOUNCE"], lanes["FEELER"], lanes["AIM"], lanes["SCATTER"], lanes["QUERY"])
        ref = ref or per_ray
        assert per_ray == ref, "queries / resolves / scatters per ray do not depend on the rule"
        assert out.busy > 0 and out.makespan * 4 >= out.busy
