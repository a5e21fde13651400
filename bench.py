#!/usr/bin/env python
"""bench.py — probe-rays/s of the DDGI probe update on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One step = one frame's probe update over the whole probe field: 4 dynamic lights are
moved (host), every probe ray is traced (ddgi_probe_update) and, on N > 1 GPUs, the
shards are exchanged (default: the kernel stores its texels into every replica over NVLink and a
one-warp epoch-flag kernel is the completion barrier; --exchange nccl: the engine's in-place
ncclAllGather, ddgi_exchange_allgather; the fused run also times the NCCL exchange and prints it
as `exchange_nccl`).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field; --verify (N > 1)
checks every replica against a full single-GPU update.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "probe_rays_per_s"
UNIT = "probe-rays/s"


# ----------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class DevPtr:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, cfg_name):
    """The reference's CPU implementation of the path = the oracle port (the GLSL cannot be
    built here or on the box: no glslang / Vulkan / lavapipe), all host threads, on a
    bounded sample of the same workload: `sample_rows` probe rows per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bench_support import reference_arm

    print(json.dumps(reference_arm(args, cfg_name)), flush=True)


# ----------------------------------------------------------------------------- live ncu pass
NCU_METRICS = ("dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,"
               "smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum")


def traffic_probe(args):
    """`bench.py --traffic-probe`: the workload's probe update a few times and nothing else - what the ncu pass
    below profiles (no torch: the engine alone)."""
    import importlib

    import ddgi_b200
    from bench_support import workload_config

    configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
    cfg = workload_config(args.workload)
    r = ddgi_b200.RVPT(*cfg["screen"])
    configs.apply(r, cfg, time=0.0)
    r.generate_probe_rays(reseed=True)
    r.set_kernel_variant(args.variant)
    r.update(advance_time=False)
    for i in range(5):
        r.render_settings.time = 2.0 * (i + 1)
        r.lights = configs.lights_for(cfg, r.render_settings.time)
        r.update(advance_time=False)
        r.probe_update()
    r.sync()
    r.close()


def ncu_pass(args, n_rays):
    """DRAM traffic, issue-slot utilisation and active lanes per instruction of ONE launch of the probe-update kernel,
    measured now by running this script's --traffic-probe mode under ncu (the 5th launch: schedule calibrated, caches
    warm as in the timed loop).  Profiler numbers explain the kernel; they are never the bench value.  None if ncu is
    missing or fails."""
    import csv
    import io
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    kernel = "probe_update_wavefront" if args.variant != 0 else "probe_update_direct"
    cmd = [ncu, "--metrics", NCU_METRICS, "--clock-control", "none", "-k", f"regex:{kernel}", "-s", "4", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--traffic-probe", "--workload", args.workload, "--variant", str(args.variant)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0")))
    except Exception:
        return None
    vals = {}
    rows = [l for l in p.stdout.splitlines() if l.startswith('"')]
    for row in csv.reader(io.StringIO("\n".join(rows))):
        if len(row) >= 3 and row[-3] in NCU_METRICS:
            try:
                v = float(row[-1].replace(",", ""))
            except ValueError:
                continue
            unit = row[-2]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,   # bytes; durations in ms
                    "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
            vals[row[-3]] = v * mult.get(unit, 1.0)
    if "dram__bytes_read.sum" not in vals:
        return None
    inst = vals.get("smsp__inst_executed.sum")
    return {"dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals.get("dram__bytes_write.sum", 0.0),
            "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_inst": vals.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warp_inst_per_ray": inst / n_rays if inst else None,
            "kernel_ms_under_ncu": vals.get("gpu__time_duration.sum"),
            "how": "ncu --metrics ... -k regex:" + kernel + " -s 4 -c 1 on `bench.py --traffic-probe` (this run)"}


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="field_32")
    ap.add_argument("--exchange", default="fused", choices=["nccl", "fused"],
                    help="N > 1: fused = the kernel stores texels into every replica over NVLink (double-buffered replicas), then a "
                         "device-side epoch-flag barrier (no collective library on the data path); nccl = the engine's in-place "
                         "all-gather (ddgi_exchange_allgather)")
    ap.add_argument("--sharding", default=None, choices=["cyclic", "slab"],
                    help="N > 1: cyclic ownership (single probes with the fused exchange, blocks of probe rows with nccl) or one "
                         "contiguous slab of probe rows per rank (default: cyclic for fused, slab for nccl)")
    ap.add_argument("--variant", type=int, default=2)
    ap.add_argument("--march-min", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--frames-in-flight", type=int, default=None, choices=[1, 2],
                    help="2 (default for --gpus > 1): consecutive updates overlap on the engine's own two streams "
                         "(ddgi_set_frames_in_flight), which hides the drain of the persistent kernel on a 1/N share; "
                         "1 (default on one GPU, where the drain is 3 %% of the kernel: profiles/r2_ab.md f)")
    ap.add_argument("--no-ncu", action="store_true", help="skip the live ncu pass (roofline.traffic falls back to profiles/traffic.json)")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--verify", action="store_true",
                    help="N > 1: after the timed run, check that every rank's replica of both texture planes equals a full "
                         "single-GPU update of the same frame (untimed)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.traffic_probe:
        return traffic_probe(args)

    if args.impl == "reference":
        run_reference(args, args.workload)
        return

    import torch
    import torch.distributed as dist

    import ddgi_b200
    from bench_support import cpu_baseline, workload_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = workload_config(args.workload)
    import importlib

    configs = importlib.import_module(ddgi_b200._pkg.__name__ + ".configs")
    fused = args.exchange == "fused"
    sharding = args.sharding or ("cyclic" if fused else "slab")

    r = ddgi_b200.RVPT(*cfg["screen"], device=local)
    configs.apply(r, cfg, time=0.0)
    r.generate_probe_rays(reseed=True)
    r.set_kernel_variant(args.variant)
    if args.march_min is not None:
        r.set_tuning(args.march_min)
    r.update(advance_time=False)
    stream = torch.cuda.current_stream()
    r.stream = stream.cuda_stream
    # updates alternate between two texture allocations: frame i can be rendered / copied out while frame i+1 is
    # traced - on one GPU and, with both allocations mapped by the peers, under the fused exchange
    r.set_double_buffer(True)

    X, Y, Z = cfg["probe_count"]
    rx, ry = cfg["tile"]
    n_rays = X * Y * Z * rx * ry
    W, H = r.probe_texture_size
    nbytes = W * H * 4
    sh = ddgi_b200.sharding
    n_probes = X * Y * Z
    row_bytes = W * 4 * ry
    slab = sh.probe_row_shard(Y, rank, world)           # the probe rows a rank reads back (e2e) in any mode

    def set_ownership(kind):
        """-> (rank of every probe, description)"""
        owner = np.zeros(n_probes, dtype=np.int32)
        if world == 1:
            return owner, "none"
        if kind == "probes":
            # blocks of 5 probes: dealt one by one, rank r of 8 would hold the probe planes x = r, r+8, r+16, r+24 of
            # the 32-wide lattice - not a uniform sample of the field (kernel time max / mean over the ranks 1.036
            # against 1.022, profiles/diag_balance.py)
            r.set_probes_cyclic(rank, world, 5)
            return ((np.arange(n_probes) // 5) % world).astype(np.int32), f"{n_probes} probes dealt round-robin in blocks of 5 to {world} ranks"
        if kind == "rows":
            block = sh.cyclic_block(Y, world)
            r.set_probe_rows_cyclic(rank, world, block)
            return np.repeat((np.arange(Y) // block) % world, X * Z).astype(np.int32), f"probe rows {Y}/{world}, block-cyclic (block {block})"
        r.set_probe_rows(*slab)
        for g in range(world):
            a, b = sh.probe_row_shard(Y, g, world)
            owner[a * X * Z:b * X * Z] = g
        return owner, f"probe rows {Y}/{world}, contiguous slabs"

    probe_owner, shard_desc = set_ownership("slab" if sharding == "slab" else "probes")
    owned_probes = probe_owner == rank

    def join_nccl():
        ids = [ddgi_b200.RVPT.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        r.comm_init(ids[0], rank, world)

    if world > 1 and fused:
        handles = [None] * world
        dist.all_gather_object(handles, r.export_texture_handle())   # both allocations of the double-buffered context
        r.open_peers(handles, rank)
    elif world > 1:
        join_nccl()

    def exchange():
        if world == 1:
            return
        if fused:
            # texels were stored into every replica by the kernel; the completion barrier is a
            # one-warp kernel exchanging epoch flags through the same peer mappings
            r.exchange_barrier()
        else:
            r.exchange_allgather()   # ncclAllGather (slabs) / grouped ncclBroadcast (block-cyclic rows), in place

    frame_no = [0]

    def prep():
        # update_lights: 4 dynamic lights, time += 2 per frame (rvpt.cpp:281); host work only
        frame_no[0] += 1
        r.render_settings.time = 2.0 * frame_no[0]
        r.lights = configs.lights_for(cfg, r.render_settings.time)
        r.update(advance_time=False)

    def step():
        prep()
        r.probe_update()
        exchange()

    # the L2 flush: one write of a buffer 1.25 x the L2 (rounded up to 32 MiB: 160 MiB for B200's 126 MB)
    l2_bytes = int(torch.cuda.get_device_properties(local).L2_cache_size)
    flush_mib = max(64, -(-(l2_bytes * 5 // 4) // (32 << 20)) * 32)
    flush_buf = None if args.no_flush else torch.empty(flush_mib << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def flush_l2():
        if flush_buf is not None:
            flush_buf.fill_(frame_no[0] & 255)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timed_launches = [0]   # kernels of ours launched inside the last timed region (this rank)

    def timed(n_steps, warm):
        """-> (ms per step, kernel ms per step), device-timed, max over ranks, L2 flushed between steps"""
        for _ in range(warm):
            flush_l2()
            step()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
                torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        timed_launches[0] = -r.launch_count
        for a, m, b in evs:
            flush_l2()
            a.record(stream)
            prep()
            r.probe_update()
            m.record(stream)
            exchange()
            b.record(stream)
        timed_launches[0] += r.launch_count
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, m, b in evs) / n_steps, sum(a.elapsed_time(m) for a, m, b in evs) / n_steps],
                         device=f"cuda:{local}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    side = torch.cuda.Stream(device=local)   # the L2 flush of the two-frames-in-flight loop

    def timed_in_flight(n_steps, warm):
        """-> ms per step with two frames in flight (ddgi_set_frames_in_flight): the updates run on the engine's own
        two streams, so per-step events on this stream would bracket nothing; K steps between two events, the second
        behind a fence on every frame in flight.  Consecutive updates overlap in time and share the L2 by
        construction, so "cold L2 per step" has no meaning here; the same write (1.25 x the L2) is still issued once per step,
        on a side stream and inside the timed region (ordered in front of an update it would hold the update back
        until the previous one has drained - the overlap this mode exists for)."""
        def run(n):
            for _ in range(n):
                if flush_buf is not None:
                    with torch.cuda.stream(side):
                        flush_buf.fill_(frame_no[0] & 255)
                step()
            r.frame_fence()
            stream.wait_stream(side)
        run(warm)
        barrier()
        if n_steps == 0:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timed_launches[0] = -r.launch_count
        a.record(stream)
        run(n_steps)
        b.record(stream)
        timed_launches[0] += r.launch_count
        barrier()
        t = torch.tensor([a.elapsed_time(b) / n_steps], device=f"cuda:{local}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # Frames in flight: on N > 1 GPUs both pipelines are measured with the same K / W and the faster one is the
    # line's `value` (config.frames_in_flight says which; both are printed).  On one GPU the drain two frames in
    # flight hide is 1 % of the kernel: one frame at a time unless asked for.
    try_two = args.frames_in_flight == 2 or (args.frames_in_flight is None and world > 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_serial, kernel_ms = timed(args.steps, args.warmup)
    launches = timed_launches[0]
    ms_two = None
    if try_two:
        r.set_frames_in_flight(2)
        ms_two = timed_in_flight(args.steps, args.warmup)
        launches_two = timed_launches[0]
    clocks = sampler.stop() if rank == 0 else None
    in_flight = ms_two is not None and (ms_two < ms_serial or args.frames_in_flight == 2)
    ms_per_step = ms_two if in_flight else ms_serial
    if in_flight:
        launches = launches_two
    else:
        r.set_frames_in_flight(1)
    args.frames_in_flight = 2 if in_flight else 1
    value = n_rays / (ms_per_step * 1e-3)

    # ---- FPS at the config's resolution: probe update + exchange + pixel pass.  On N > 1 GPUs every
    #      replica holds the whole texture after the exchange, so each rank renders 1/N of the frame's
    #      16-pixel workgroup rows (no further collective; the bands are read back independently) ----
    band = r.set_frame_band(rank, world)
    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_frames = max(3, min(args.steps, 10))
    barrier()
    fa.record(stream)
    for _ in range(n_frames):
        step()
        r.render_frame()
    r.frame_fence()
    fb.record(stream)
    barrier()
    ft = torch.tensor([fa.elapsed_time(fb) / n_frames], device=f"cuda:{local}", dtype=torch.float64)
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r.frame_fence()
    pa.record(stream)
    r.render_frame()
    pb.record(stream)
    barrier()
    if world > 1:
        dist.all_reduce(ft, op=dist.ReduceOp.MAX)
    frame_ms = float(ft[0])
    pixel_ms = pa.elapsed_time(pb)

    # ---- algorithmic bytes: the voxel lookups of the REFERENCE ALGORITHM per ray, counted by the instrumented kernel in
    #      variant 1 (every lookup the reference performs; the default variant 2 ends shadow feelers early and performs
    #      fewer - SURVEY 8d: result-preserving early-outs do not reduce the algorithmic figure).  Untimed. ----
    r.set_debug(True)
    r.set_kernel_variant(1 if args.variant == 2 else args.variant)
    r.probe_update()
    exchange()
    r.sync()
    lk_all = r.read_lookup_counts(0).reshape(n_probes, rx * ry)
    lk_sum = torch.tensor([float(lk_all[owned_probes].sum(dtype=np.float64))], device=f"cuda:{local}", dtype=torch.float64)
    performed = None
    if args.variant == 2:
        r.set_kernel_variant(2)
        r.probe_update()
        exchange()
        r.sync()
        performed = torch.tensor([float(r.read_lookup_counts(0).reshape(n_probes, rx * ry)[owned_probes].sum(dtype=np.float64))],
                                 device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(lk_sum)
        if performed is not None:
            dist.all_reduce(performed)
    mean_lookups = float(lk_sum[0]) / n_rays
    # the pixel pass against SURVEY 8d's bytes per pixel: 4 B x voxel lookups (primary + feelers, counted by the
    # instrumented kernel) + 8 probes x (1 + taps) x 4 B + the 4-B store, taps at their upper bound of 25
    pixel_roofline = None
    if world == 1:
        r.render_frame()
        r.sync()
        w_px, h_px = cfg["screen"]
        px_lk = r.read_lookup_counts(1).reshape(h_px, w_px)[band[0]:band[1], :(w_px // 16) * 16]
        bytes_px = 4.0 * float(px_lk.mean()) + 8 * 26 * 4 + 4
        px_gbs = px_lk.size * bytes_px / (pixel_ms * 1e-3) / 1e9
        pixel_roofline = {"bound": "hbm", "kernel": "render_frame_kernel", "kernel_ms": pixel_ms, "pixels": int(px_lk.size),
                          "mean_lookups_per_pixel": float(px_lk.mean()), "bytes_per_pixel": bytes_px, "achieved": px_gbs,
                          "unit": "GB/s", "note": "SURVEY 8d bytes per pixel with the tile gathers at their upper bound (25 taps)"}
    r.set_kernel_variant(args.variant)
    r.set_debug(False)
    bytes_per_ray = 4.0 * mean_lookups + 8.0
    peak, peak_src = load_peaks()
    rays_this_rank = int(owned_probes.sum()) * rx * ry
    achieved = rays_this_rank * bytes_per_ray / (kernel_ms * 1e-3) / 1e9

    # ---- e2e through the C-ABI with host buffers (pinned), per step:
    #      H2D: ray-sample table + uniforms/lights; D2H: the rank's share of the albedo plane.
    #      The texture is double-buffered in the engine (under the fused exchange both allocations are mapped by the
    #      peers), so the D2H of step i (asynchronous, on the engine's copy stream, into one of two pinned buffers)
    #      overlaps the trace of step i+1 at every N; every step's rows are still read back inside the timed region ----
    e2e = None
    if not args.no_e2e:
        samples = r.ray_samples       # the stratified sample table generate_samples drew
        pinned_samples = torch.from_numpy(samples).pin_memory()
        lib = ddgi_b200.capi.load()
        h2d = pinned_samples.numel() * 4 + 32 + 48 + 80 + 4 * 28
        rows = (0, H) if world == 1 else (slab[0] * ry, slab[1] * ry)
        d2h = (rows[1] - rows[0]) * W * 4
        host_pair = [torch.empty(max(d2h, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]

        def e2e_step():
            prep()   # host side; overlaps the previous step's trace (set_ray_samples then waits for that trace)
            rc = lib.ddgi_set_ray_samples(r._ctx, pinned_samples.data_ptr(), rx * ry)
            assert rc == 0
            r.probe_update()
            exchange()
            r.read_probe_texture_rows_async(host_pair[frame_no[0] & 1].data_ptr(), rows[0], rows[1], d2h, 0)

        def e2e_run():
            """-> whole-job probe-rays/s, wall clock around K steps including every copy, max over ranks"""
            for _ in range(3):
                e2e_step()
            r.read_wait()
            r.sync()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            r.read_wait()
            r.sync()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local}", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return n_rays * args.steps / float(dt[0])

        # both pipelines at N > 1 (as for `value`), the faster one is the line's e2e
        e2e_by = {args.frames_in_flight: e2e_run()}
        if try_two:
            other = 3 - args.frames_in_flight
            r.set_frames_in_flight(other)
            e2e_by[other] = e2e_run()
            r.set_frames_in_flight(args.frames_in_flight)
        e2e_fl = max(e2e_by, key=e2e_by.get)
        e2e = {"value": e2e_by[e2e_fl], "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "inputs": "ray-sample table + uniforms + lights (pinned host)",
               "result": "albedo probe texture rows of this rank (pinned host)",
               "frames_in_flight": e2e_fl, "by_frames_in_flight": {str(k): v for k, v in sorted(e2e_by.items())},
               "pipelining": "double-buffered texture (replicas mapped pairwise under the fused exchange): the D2H of step i overlaps the trace of step i+1"}
        # literal storage-buffer mode: the whole ProbeRay array re-uploaded every frame, as
        # RVPT::update does (rvpt.cpp:285)
        if world == 1:
            rays_host = r.probe_rays      # the reference's std::vector<ProbeRay>
            pinned_rays = torch.from_numpy(rays_host).pin_memory()
            host_tex = torch.empty(nbytes, dtype=torch.uint8).pin_memory()

            def ssbo_step():
                frame_no[0] += 1
                r.update(advance_time=False)
                rc = lib.ddgi_set_probe_rays(r._ctx, pinned_rays.data_ptr(), n_rays)
                assert rc == 0
                r.probe_update()
                rc = lib.ddgi_read_probe_texture(r._ctx, 0, 0, host_tex.data_ptr(), nbytes)
                assert rc == 0

            for _ in range(2):
                ssbo_step()
            barrier()
            t0 = time.perf_counter()
            ns = max(3, args.steps // 4)
            for _ in range(ns):
                ssbo_step()
            barrier()
            e2e["ssbo_mode"] = {"value": n_rays * ns / (time.perf_counter() - t0), "unit": UNIT,
                                "h2d_bytes_per_step": int(n_rays * 48 + 160), "d2h_bytes_per_step": int(nbytes)}
            r.generate_probe_rays(reseed=True)

    def replica():
        ptr, nb = r.probe_texture_device_ptr(0)   # (the current allocation: it alternates)
        return torch.as_tensor(DevPtr(ptr, 2 * nb), device=f"cuda:{local}")

    # ---- untimed: the exchanged replicas against a full local update of the same frame ----
    verify = None
    if args.verify and world > 1:
        step()
        barrier()
        got = replica().clone()
        if fused:
            r.exchange_status()   # raises if any barrier timed out
            r.close_peers()
            fused = False
        r.set_probe_rows(0, Y)    # this rank alone, every probe, the same frame (lights unchanged since step())
        r.probe_update()
        barrier()
        ok = torch.tensor([1 if torch.equal(got, replica()) else 0], device=f"cuda:{local}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verify = {"replicas_equal_full_update": bool(int(ok[0])), "bytes_compared_per_rank": int(got.numel())}

    # ---- N > 1, fused default: the same workload once more with the engine's NCCL exchange (ddgi_exchange_allgather),
    #      device-timed like the headline: (a) the balanced probe ownership of the headline, each rank's tiles packed
    #      into one chunk, ONE ncclAllGather, unpacked; (b) contiguous slabs of probe rows, ONE in-place ncclAllGather
    #      per plane and no packing (SURVEY 8e's layout; the slabs are not equally expensive) ----
    exchange_nccl = None
    if world > 1 and args.exchange == "fused":
        if fused:
            r.exchange_status()
            barrier()
            r.close_peers()
            fused = False
        join_nccl()
        n_nccl = max(3, min(args.steps, 10))
        r.set_frames_in_flight(1)
        set_ownership("probes")
        ms_nccl, kernel_ms_nccl = timed(n_nccl, 3)
        exchange_nccl = {"value": n_rays / (ms_nccl * 1e-3), "unit": UNIT, "ms_per_step": ms_nccl, "kernel_ms": kernel_ms_nccl,
                         "steps": n_nccl, "exchange": "nccl: probes dealt round-robin in blocks of 5 (the headline's ownership), the rank's "
                                                      "tiles packed into one chunk, ONE ncclAllGather, unpacked (ddgi_exchange_allgather; the "
                                                      "distance plane only holds zeros and is skipped)"}
        if Y % world == 0:
            set_ownership("slab")
            ms_slab, kernel_ms_slab = timed(n_nccl, 3)
            exchange_nccl["slabs"] = {"value": n_rays / (ms_slab * 1e-3), "unit": UNIT, "ms_per_step": ms_slab, "kernel_ms": kernel_ms_slab,
                                      "exchange": "nccl: contiguous slabs of probe rows, one in-place ncclAllGather per plane, no packing"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r.set_frames_in_flight(1)
        r.set_double_buffer(False)
        cpu = cpu_baseline(r, cfg, args.workload)
    prof = None
    if rank == 0 and world == 1 and not args.no_ncu:
        prof = ncu_pass(args, n_rays)

    if rank == 0:
        traffic, traffic_src = None, None
        if prof:
            traffic, traffic_src = prof["dram_bytes_per_launch"], prof["how"]
        elif world == 1:
            traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(traffic_file):
                try:
                    with open(traffic_file) as f:
                        tj = json.load(f)
                    if tj.get("workload") == args.workload:
                        traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/traffic.json (committed ncu capture, not this run)"
                except Exception:
                    pass
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "probes": [X, Y, Z], "rays_per_probe": rx * ry,
                       "probe_rays": n_rays, "voxels": list(cfg["voxels"][1]), "lights": 4 if cfg["lights"] == "cave4" else 1,
                       "max_bounces": cfg.get("max_bounces", 8), "resolution": list(cfg["screen"]),
                       "l2": ("not flushed" if flush_buf is None else
                              f"one {flush_mib} MiB write (L2 = {l2_bytes / 1e6:.0f} MB) per step on a side stream, inside the timed region (two frames in flight share the L2)" if in_flight else
                              f"flushed between timed steps ({flush_mib} MiB write, L2 = {l2_bytes / 1e6:.0f} MB)"),
                       "frames_in_flight": args.frames_in_flight,
                       "kernel_variant": args.variant, "exchange": args.exchange if world > 1 else "none",
                       "sharding": shard_desc if world > 1 else "none"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "probe_update_wavefront" if args.variant != 0 else "probe_update_direct",
                         "kernel_ms": kernel_ms, "bytes_per_ray": bytes_per_ray, "mean_lookups_per_ray": mean_lookups,
                         "lookups_performed_per_ray": (float(performed[0]) / n_rays) if performed is not None else mean_lookups,
                         "note": "algorithmic bytes = rays x (4 B x voxel lookups of the reference algorithm + 8 B texel stores), SURVEY 8d; "
                                 "the kernel is bound by instruction issue, not by HBM: see `issue`"},
            "issue": None if not prof else {k: prof[k] for k in ("issue_active_pct", "lanes_per_inst", "warp_inst_per_ray", "kernel_ms_under_ncu", "how")},
            "pipelines": {"one_frame_at_a_time": {"value": n_rays / (ms_serial * 1e-3), "unit": UNIT, "ms_per_step": ms_serial, "kernel_ms": kernel_ms,
                                                  "note": "ddgi_set_frames_in_flight(1): per-step CUDA events, L2 flushed between steps; roofline.kernel_ms is this pass's"},
                          "two_frames_in_flight": None if ms_two is None else
                          {"value": n_rays / (ms_two * 1e-3), "unit": UNIT, "ms_per_step": ms_two,
                           "note": "ddgi_set_frames_in_flight(2): K steps between two events behind a fence on every frame in flight"},
                          "value_is": "two_frames_in_flight" if in_flight else "one_frame_at_a_time"},
            "cpu_baseline": cpu,
            "verify": verify,
            "exchange_nccl": exchange_nccl,
            "fps": {"value": 1000.0 / frame_ms, "frame_ms": frame_ms, "pixel_pass_ms": pixel_ms,
                    "resolution": list(cfg["screen"]), "pixel_rows_rank0": list(band),
                    "pixel_roofline": None if not pixel_roofline else dict(pixel_roofline, peak=peak, frac=pixel_roofline["achieved"] / peak),
                    "note": "probe update + exchange + pixel pass; pixel rows split across ranks"},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        if fused:
            r.exchange_status()  # raises if any barrier timed out
            r.close_peers()
        dist.barrier()
        dist.destroy_process_group()
    r.close()


if __name__ == "__main__":
    main()
