"""Host model of the experimental pooled kernel (variant 2): issue counts of its block logic on field_8 for a few
pool sizes, with the block's warps taking turns or all holding their rays while the others claim, and the
modelled warp instructions per ray (profiles/r1_policy_model.md).  No GPU."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import util
from oracle import oracle
hs = util.hostsim()
hs.sim_probe_update_pooled_stats.restype = C.c_uint64
hs.sim_probe_update_pooled_stats.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
cfg = util.configs.CONFIGS["field_8"]
sc = util.oracle_scene(cfg)
rx, ry = cfg["tile"]
rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True)))
n = rays.shape[0]
W, H = sc.tex_size
# calibrated per-issue code costs of variant 1 (profiles/policy_sim.py): MARCH step 76 (incl. loop ballot), QUERY 98, BOUNCE 299, FEELER 129, SCATTER 303, FETCH 209
Q = ["MARCH", "BOUNCE", "FEELER", "FETCH", "SLOW"]
for blocks, keep, slots, lock in ((16, 16, 128, 0), (16, 16, 128, 1), (16, 24, 128, 1), (16, 16, 256, 1), (16, 24, 256, 1), (16, 16, 384, 1), (16, 24, 384, 1), (16, 24, 512, 1)):
    st = np.zeros(16, dtype=np.uint64)
    alb = np.zeros((H, W), dtype=np.uint32)
    hs.sim_probe_update_pooled_stats(C.byref(sc.p), rays.ctypes.data, n, blocks, keep, alb.ctypes.data, None, st.ctypes.data, slots, lock)
    issues = st[:5].astype(float); lanes = st[5:10].astype(float)
    march_iters, march_lanes = float(st[10]), float(st[11])
    pick_claim_push = 25 + 25 + 20     # pick a queue, claim with CAS, push to the new queues (estimates)
    sess = issues[0] + issues[4]
    cost = (march_iters * 76 + sess * (7 * 2 + pick_claim_push)                     # march: 4 loads + 3 stores (x2: address + op)
            + issues[1] * (299 + 98 + 18 * 2 + pick_claim_push)                        # bounce: resolve + query, 9 + 9 vectors
            + issues[2] * (129 + 0.55 * 303 + 98 + 18 * 2 + pick_claim_push)           # feeler: resolve (+ scatter about half the time) + query
            + issues[3] * (209 + 98 + 13 * 2 + pick_claim_push))
    print(f"slots {slots} lockstep {lock} keep {keep:2d}: issues/ray march-sessions {sess / n:.2f} (iters {march_iters / n:.2f}, {march_lanes / max(march_iters, 1):.1f} lanes/iter, "
          f"{lanes[0] / max(issues[0], 1):.1f} claimed), bounce {issues[1] / n:.3f} ({lanes[1] / max(issues[1], 1):.1f} lanes), feeler {issues[2] / n:.3f} "
          f"({lanes[2] / max(issues[2], 1):.1f}), fetch {issues[3] / n:.3f} ({lanes[3] / max(issues[3], 1):.1f}); idle passes/ray {st[12] / n:.2f}; "
          f"modelled warp-inst/ray {cost / n:.0f} (variant 1: 767)")
