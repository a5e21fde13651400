#!/usr/bin/env python
"""Aggregates an ncu source-page CSV of one kernel by source function.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c ddgi_kernels.sm_100a.cubin > k.dis
    python profiles/prof_by_function.py src.csv k.dis probe_update_wavefront

The n-th SASS instruction of the kernel in k.dis (which carries //## File/line markers)
is the n-th row of the CSV; each row is attributed to the innermost source line and that
line to the function whose definition precedes it in its file.
"""
import csv
import re
import sys
from collections import defaultdict


def functions_of(path):
    starts = []
    try:
        for n, line in enumerate(open(path), 1):
            m = re.match(r"^(?:DDGI_HD|DDGI_D|__global__|__device__|static|template|cudaError_t)[^;]*?\b([A-Za-z_]\w*)\s*\(", line)
            if m and not line.rstrip().endswith(";"):
                starts.append((n, m.group(1)))
    except OSError:
        pass
    return starts


def main():
    src_csv, dis, kernel = sys.argv[1:4]
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    # instruction -> (file, line) from nvdisasm
    locs = []
    cur = ("?", 0)
    inside = False
    for line in open(dis):
        if line.startswith(".text."):
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", line):
            locs.append(cur)
    if len(locs) != len(data):
        print(f"warning: {len(locs)} instructions in the disassembly vs {len(data)} rows in the CSV", file=sys.stderr)
    fn_cache = {}
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    tot_w = tot_t = tot_s = 0.0
    for (f, l), r in zip(locs, data):
        if f not in fn_cache:
            fn_cache[f] = functions_of(f)
        name = "?"
        for n, fn in fn_cache[f]:
            if n <= l:
                name = fn
            else:
                break
        w = float(r[ix["Instructions Executed"]] or 0)
        t = float(r[ix["Thread Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        key = f"{f.split('/')[-1]}:{name}"
        a = agg[key]
        a[0] += w
        a[1] += t
        a[2] += s
        a[3] += 1
        tot_w += w
        tot_t += t
        tot_s += s
    print(f"kernel {kernel}: {len(data)} SASS instructions, {tot_w:.4g} warp instructions executed, "
          f"{tot_t / max(tot_w, 1):.2f} threads / instruction")
    print(f"{'function':46s} {'SASS':>5s} {'warp-inst %':>11s} {'thr/inst':>8s} {'samples %':>9s}")
    for k, (w, t, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if w / tot_w < 0.002:
            continue
        print(f"{k:46s} {n:5d} {100 * w / tot_w:11.1f} {t / max(w, 1):8.1f} {100 * s / max(tot_s, 1):9.1f}")


if __name__ == "__main__":
    main()
