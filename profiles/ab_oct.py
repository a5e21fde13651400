"""A/B timing of the octahedral layout's probe update (trace into the ray buffer + probe_blend_octahedral), 1 GPU, cold L2:
    DDGI_LIB=profiles/ab/libddgi_x.so python profiles/ab_oct.py [workload=field_32] [oct=8] [n=15]
Prints the CUDA-event time of the whole update and a CRC of both planes.  Not a bench value."""
import importlib, os, sys, zlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddgi_b200
from bench_support import workload_config
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
name = sys.argv[1] if len(sys.argv) > 1 else "field_32"
oct_ = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 15
lib = os.path.basename(os.environ.get("DDGI_LIB") or "default")
cfg = workload_config(name)
r = ddgi_b200.RVPT(*cfg["screen"])
configs.apply(r, cfg)
r.set_layout(ddgi_b200.capi.LAYOUT_OCTAHEDRAL, oct_)
r.generate_fibonacci_rays()
r.update(advance_time=False)
stream = torch.cuda.current_stream(); r.stream = stream.cuda_stream
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4): r.probe_update()
torch.cuda.synchronize()
ts = []
for _ in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1); a.record(stream); r.probe_update(); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
crc = zlib.crc32(r.read_probe_texture(0).tobytes() + r.read_probe_texture(1).tobytes())
print(f"{lib:24s} {name} octahedral {oct_}x{oct_}: update (trace + blend) median {np.median(ts):.3f} ms  min {min(ts):.3f}  crc {crc:08x}", flush=True)
r.close()
