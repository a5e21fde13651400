"""A/B of the schedule granularity on field_32 (1 GPU, cold L2, median of 10): full field and a
1/8 share, slots of 32 / 8 / 4 / 1 rays.  Not a bench value."""
import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddgi_b200
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
cfg = configs.CONFIGS["field_32"]
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream(); r.stream = stream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for world in (1, 8):
    for slot in (32, 8, 4, 1):
        r.set_probes_cyclic(0, world, 1)
        r.set_schedule_slot(slot)
        for _ in range(3): r.probe_update()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1); a.record(stream); r.probe_update(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        tex = r.read_probe_texture(0)
        if world == 1:
            ref = tex.copy() if ref is None else ref
            assert np.array_equal(tex, ref), "the schedule changed a result"
        print(f"world {world} slot {slot:2d}: median {np.median(ts):.3f} ms")
