"""A/B timing of the probe-update kernel on field_32 (1 GPU, cold L2, median of 15): run with DDGI_LIB=<alternative build> to compare builds.  Not a bench value."""
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
for _ in range(4): r.probe_update()
torch.cuda.synchronize()
ts = []
for _ in range(15):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1); a.record(stream); r.probe_update(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(os.environ.get("DDGI_LIB", "default"), f"median {np.median(ts):.3f} min {min(ts):.3f}")
