"""Worker of tests/test_multi_gpu_fused.py: one process per GPU.  Rank r of `world` owns its share of field_8's probes
(round-robin in blocks of 5), updates it for `frames` frames with MOVING lights under the fused exchange (texels stored
into every peer's replica, epoch barriers), renders its band of every frame and reads the whole replica back
asynchronously every frame - the pattern in which a faster rank could store frame i+1 into a replica whose owner still
reads frame i (round 1's ADVICE).  Everything read is written to `out` for the parent to compare with a single-GPU run.
With exchange = nccl the same loop runs over ddgi_exchange_allgather instead (probe-cyclic ownership: tiles packed into one
chunk per rank, ONE ncclAllGather, unpacked), no peer mappings.
    python fused_worker.py rank world frames_in_flight double_buffer exchange_dir out.npz [fused|nccl]"""
import importlib
import os
import pickle
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddgi_b200  # noqa: E402

rank, world, in_flight, double_buffer = (int(v) for v in sys.argv[1:5])
xdir, out = sys.argv[5], sys.argv[6]
nccl = len(sys.argv) > 7 and sys.argv[7] == "nccl"
configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
cfg = configs.CONFIGS["field_8"]
FRAMES = 8


def exchange_files(tag, payload):
    """all-gather of small byte strings through files (no torch.distributed in the worker)"""
    with open(os.path.join(xdir, f"{tag}.{rank}.tmp"), "wb") as f:
        pickle.dump(payload, f)
    os.rename(os.path.join(xdir, f"{tag}.{rank}.tmp"), os.path.join(xdir, f"{tag}.{rank}"))
    got = []
    for g in range(world):
        p = os.path.join(xdir, f"{tag}.{g}")
        t0 = time.time()
        while not os.path.exists(p):
            if time.time() - t0 > 120:
                raise SystemExit(f"rank {rank}: rank {g} never wrote {tag}")
            time.sleep(0.01)
        with open(p, "rb") as f:
            got.append(pickle.load(f))
    return got


with ddgi_b200.RVPT(*cfg["screen"], device=rank) as r:
    configs.apply(r, cfg)
    r.generate_probe_rays(reseed=True)
    r.update(advance_time=False)
    if double_buffer:
        r.set_double_buffer(True)
    r.set_probes_cyclic(rank, world, 5)
    band = r.set_frame_band(rank, world)
    if nccl:
        ids = exchange_files("id", ddgi_b200.RVPT.comm_unique_id() if rank == 0 else b"")
        r.comm_init(ids[0], rank, world)
    else:
        handles = exchange_files("handle", r.export_texture_handle())
        r.open_peers(handles, rank)
    if in_flight == 2:
        r.set_frames_in_flight(2)
    W, H = r.probe_texture_size
    tex = [np.zeros((H, W), dtype=np.uint32) for _ in range(FRAMES)]
    frames = []
    for f in range(FRAMES):
        r.render_settings.time = 2.0 * (f + 1)
        r.lights = configs.lights_for(cfg, r.render_settings.time)
        r.update(advance_time=False)
        r.probe_update()
        if nccl:
            r.exchange_allgather()
        else:
            r.exchange_barrier()
        if rank == 0:
            for _ in range(40):   # rank 0 reads frame f for a long time on the device: the peers must not run ahead into its replica
                r.render_frame()
        if double_buffer:
            r.read_probe_texture_async(tex[f].ctypes.data, tex[f].nbytes, 0)
        r.render_frame()
        if not double_buffer:
            tex[f][:] = r.read_probe_texture(0)
        if rank == 0 and f % 3 == 1:
            time.sleep(0.02)   # one rank falls behind on the host now and then
        if f % 2 == 1:
            frames.append(r.read_frame()[band[0]:band[1]].copy())   # (synchronous: this frame's band)
    r.read_wait()
    r.frame_fence()
    r.sync()
    if not nccl:
        r.exchange_status()
    exchange_files("done", b"")   # nobody unmaps while a peer may still store
    r.set_frames_in_flight(1)
    if not nccl:
        r.close_peers()
    np.savez(out, tex=np.stack(tex), frames=np.stack(frames), band=np.array(band))
print("ok", rank)
