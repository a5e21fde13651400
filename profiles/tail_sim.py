"""Host model of schedule orderings (profiles/r1_policy_model.md postscript, r1_tail.md): ranks rays / slots by
lookups or by a modelled time cost and replays the warp loop of variant 1.  No GPU; the device did not
confirm the modelled gain of per-ray ranking."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests'); sys.path.insert(0, ROOT + '/profiles')
import util
from oracle import oracle
import policy_sim as ps
cfg = util.configs.CONFIGS["field_8"]
sc = util.oracle_scene(cfg)
rx, ry = cfg["tile"]
rays = np.ascontiguousarray(oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True)))
n = rays.shape[0]
hs = util.hostsim()
hs.sim_ray_profile.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_uint32, C.c_void_p]
prof = np.zeros((n, 4), dtype=np.uint32)
hs.sim_ray_profile(C.byref(sc.p), rays.ctypes.data, n, prof.ctypes.data)
lk, q, b, f = (prof[:, i].astype(np.float64) for i in range(4))
cost = np.array([76, 98, 299, 129, 303, 209, 564, 0, 23], dtype=np.float64)
time_cost = 76 * lk + 98 * q + 299 * b + 129 * f + 303 * b
print("rays", n, "mean lookups", lk.mean(), "queries", q.mean(), "bounces", b.mean(), "feelers", f.mean(), "mean time cost", time_cost.mean())
hs.sim_wavefront_policy.argtypes = [C.POINTER(oracle.OrcParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(ps.PolicyOut), C.c_int]
def order_by(metric, slot=32, red=np.max):
    m = red(metric.reshape(-1, slot), axis=1)
    so = np.argsort(-m, kind="stable")
    return np.ascontiguousarray((so[:, None] * slot + np.arange(slot)[None, :]).reshape(-1).astype(np.uint32))
def run(order, n_warps):
    out = ps.PolicyOut()
    hs.sim_wavefront_policy(C.byref(sc.p), rays.ctypes.data, order.ctypes.data, n, n_warps, 0, 16, 0, cost.ctypes.data, C.byref(out), 1)
    return out.makespan, out.busy / n_warps
for n_warps in (64, 512):
    print(f"--- {n_warps} warps = {n / (32 * n_warps):.0f} rays per lane")
    for name, order in [("natural order", np.arange(n, dtype=np.uint32)),
                        ("max lookups per 32-slot (kernel today)", order_by(lk)),
                        ("max lookups + 64 x queries", order_by(lk + 64 * q)),
                        ("max modelled time", order_by(time_cost)),
                        ("sum modelled time", order_by(time_cost, red=np.sum)),
                        ("max modelled time, 256-ray slots (probes)", order_by(time_cost, slot=256)),
                        ("per-ray max modelled time (slot 1)", order_by(time_cost, slot=1))]:
        mk, avg = run(order, n_warps)
        print(f"{name:48s} makespan {mk / 1e3:9.1f}k  mean busy {avg / 1e3:9.1f}k  ratio {mk / avg:.3f}")
print("--- slot size sweep, 64 warps (mean busy = modelled warp instructions per warp)")
for slot in (1, 2, 4, 8, 16, 32, 64):
    for name, metric in (("lookups", lk), ("lookups + 64 q", lk + 64 * q)):
        mk, avg = run(order_by(metric, slot=slot), 64)
        print(f"slot {slot:3d} by max {name:16s} mean busy {avg / 1e3:9.1f}k  makespan {mk / 1e3:9.1f}k")
