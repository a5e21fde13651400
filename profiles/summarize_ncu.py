#!/usr/bin/env python
"""Writes the judged summary of one `ncu --set full` capture (read here, no GPU needed).

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep probe_update_wavefront field_32 8388608 > profiles/<name>.txt
    python profiles/summarize_ncu.py gpurun_out/prof_px.ncu-rep render_frame_kernel field_32 2058240 > profiles/<name>.txt   (pixels)

Also refreshes profiles/traffic.json (dram bytes per launch; bench.py copies it into
roofline.traffic for the same workload).
"""
import csv
import io
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * mult.get(unit, 1)


def main():
    rep, kernel, workload, n_rays = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, d = rows[0], rows[1], rows[2]
    m = {n: (d[i], u[i]) for i, n in enumerate(h)}
    print(f"# ncu --set full --clock-control none --import-source on, kernel {m['Kernel Name'][0][:70]}")
    unit_name = "probe ray" if kernel.startswith("probe_update") else "pixel"
    print(f"# capture {os.path.basename(rep)}; workload {workload}; {int(n_rays)} {unit_name}s per launch")
    for k in KEYS:
        if k in m:
            print(f"{k:70s} {m[k][0]:>18s} {m[k][1]}")
    rd = to_bytes(*m["dram__bytes_read.sum"])
    wr = to_bytes(*m["dram__bytes_write.sum"])
    print(f"{'dram bytes per launch (read + write)':70s} {rd + wr:18.0f} byte  = {(rd + wr) / n_rays:.1f} B / {unit_name}")
    if not kernel.startswith("probe_update"):
        # the pixel pass: no state machine to split by; the per-function split of the source page instead
        lib = os.path.join(ROOT, "dynamic-diffuse-global-illumination-minecraft_b200", "libddgi_b200.so")
        return
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump({"workload": workload, "kernel": kernel, "capture": os.path.basename(rep),
                   "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr}, f, indent=1)
    # per-state split of the source page
    lib = os.path.join(ROOT, "dynamic-diffuse-global-illumination-minecraft_b200", "libddgi_b200.so")
    with tempfile.TemporaryDirectory() as t:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(t, "src.csv"), "w").write(src)
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=t, capture_output=True)
        dis = subprocess.run(["nvdisasm", "-gi", "-c", "ddgi_kernels.sm_100a.cubin"], cwd=t, capture_output=True, text=True).stdout
        open(os.path.join(t, "k.dis"), "w").write(dis)
        print()
        print("# split by state of the per-lane state machine (profiles/prof_by_state.py; the library must be the profiled build)")
        sys.stdout.flush()
        # the palette instantiation <false> is the one the bench runs: select it by its mangled name
        mangled = kernel + "ILb0ELb0ELb0E" if kernel == "probe_update_wavefront" else kernel
        subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "prof_by_state.py"), os.path.join(t, "src.csv"),
                        os.path.join(t, "k.dis"), mangled, str(n_rays)])


if __name__ == "__main__":
    main()
