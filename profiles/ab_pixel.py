"""A/B timing of the pixel pass (1 GPU, median of N launches at the workload's resolution): run with DDGI_LIB=<alternative
build> to compare builds on the same box.  One line per workload with a CRC of the frame.  Not a bench value.

    DDGI_LIB=profiles/ab/libddgi_x.so python profiles/ab_pixel.py field_32[,cave_128,...] [n=15]
"""
import importlib, os, sys, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddgi_b200
from bench_support import workload_config
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
names = (sys.argv[1] if len(sys.argv) > 1 else "field_32").split(",")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
lib = os.path.basename(os.environ.get("DDGI_LIB", "default"))
for name in names:
    cfg = workload_config(name)
    r = ddgi_b200.RVPT(*cfg["screen"])
    configs.apply(r, cfg)
    r.generate_probe_rays(reseed=True)
    r.update(advance_time=False)
    stream = torch.cuda.current_stream(); r.stream = stream.cuda_stream
    r.probe_update(); r.probe_update()
    for _ in range(3): r.render_frame()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); r.render_frame(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    crc = zlib.crc32(r.read_frame().tobytes())
    w, h = cfg["screen"]
    print(f"{lib:28s} {name:12s} pixel pass {w}x{h}: median {np.median(ts):.3f} ms  min {min(ts):.3f}  crc {crc:08x}", flush=True)
    r.close()
