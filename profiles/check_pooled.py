"""EXPERIMENTAL variant 2 (pooled kernel): equality with variant 1 on two small fields, then kernel
time on field_32 for a few march_keep settings.  Run under `timeout` — a persistent kernel that
loses a ray would never end.  Not a bench value."""
import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddgi_b200
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
for name in ("cornell_3x3x3", "field_8"):
    cfg = dict(configs.CONFIGS[name]); cfg["screen"] = (64, 64)
    with ddgi_b200.RVPT(64, 64) as r:
        r.set_debug(True)
        configs.apply(r, cfg)
        r.generate_probe_rays(reseed=True)
        r.update(advance_time=False)
        out = {}
        for variant in (1, 2):
            r.set_kernel_variant(variant)
            r.write_probe_texture(np.zeros(r.probe_texture_size[::-1], dtype=np.uint32))
            for _ in range(2):
                r.probe_update()
            r.sync()
            out[variant] = (r.read_probe_texture(0).copy(), r.read_lookup_counts(0).copy(), r.read_probe_texture(0, ddgi_b200.capi.FMT_F32).copy())
        same = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(out[1], out[2]))
        print(f"{name}: variant 2 == variant 1: {same}", flush=True)
        if not same:
            sys.exit(1)
cfg = configs.CONFIGS["field_32"]
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.generate_probe_rays(reseed=True)
r.update(advance_time=False)
stream = torch.cuda.current_stream(); r.stream = stream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for variant, keep in ((1, 16), (2, 16), (2, 12), (2, 24), (2, 8)):
    r.set_kernel_variant(variant)
    r.set_tuning(keep)
    for _ in range(3): r.probe_update()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1); a.record(stream); r.probe_update(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    tex = r.read_probe_texture(0)
    ref = tex.copy() if ref is None else ref
    print(f"field_32 variant {variant} march_min/keep {keep}: median {np.median(ts):.3f} ms, texture equal to variant 1: {np.array_equal(tex, ref)}", flush=True)
