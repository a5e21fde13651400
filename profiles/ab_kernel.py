"""A/B timing of the probe-update kernel (1 GPU, cold L2, median of N launches): run with DDGI_LIB=<alternative build> to
compare builds on the same box.  Prints one line per (workload, variant, march_min) with a CRC of the albedo plane so that
builds can be checked for identical output.  Not a bench value.

    DDGI_LIB=profiles/ab/libddgi_x.so python profiles/ab_kernel.py field_32[,cave_128,...] [variants=2] [march_mins=16] [n=15]
"""
import importlib, os, sys, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddgi_b200
from bench_support import workload_config
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
names = (sys.argv[1] if len(sys.argv) > 1 else "field_32").split(",")
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2").split(",")]
mms = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "16").split(",")]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 15
lib = os.path.basename(os.environ.get("DDGI_LIB") or "default")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in names:
    cfg = workload_config(name)
    r = ddgi_b200.RVPT(*cfg["screen"])
    configs.apply(r, cfg)
    r.generate_probe_rays(reseed=True)
    r.update(advance_time=False)
    if os.environ.get('SLOT'): r.set_schedule_slot(int(os.environ['SLOT']))
    stream = torch.cuda.current_stream(); r.stream = stream.cuda_stream
    rays = r.num_probe_rays
    for variant in variants:
        try:
            r.set_kernel_variant(variant)
        except Exception as e:
            print(f"{lib} {name} variant {variant}: {e}"); continue
        for mm in mms:
            r.set_tuning(mm)
            for _ in range(4): r.probe_update()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.fill_(1); a.record(stream); r.probe_update(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            crc = zlib.crc32(r.read_probe_texture(0).tobytes())
            med = float(np.median(ts))
            print(f"{lib:28s} {name:12s} variant {variant} march_min {mm:2d}: median {med:.3f} ms  min {min(ts):.3f}  {rays / med / 1e6:8.1f} M probe-rays/s  crc {crc:08x}", flush=True)
    r.close()
