#!/usr/bin/env python
"""Dynamic profile of one kernel by source line: joins the ncu source page (instructions executed, stall
samples per SASS instruction) with the inline chains of `nvdisasm -gi` for the same build, and sums per
(kernel line, first inlined line below it).  Shows where the executed instructions of each state go.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    python profiles/prof_by_line.py src.csv <library.so> [kernel-substring] [rays] [depth]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def chains(lib, want):
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if f.startswith("ddgi_kernels.")][0]
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cub)], capture_output=True, text=True).stdout
    frame = re.compile(r'File "([^"]+)", line (\d+)')
    out, pending, annot, in_fn = [], [], False, False
    for line in txt.splitlines():
        if line.startswith(".text."):
            in_fn = want in line
            continue
        if not in_fn:
            continue
        if "//## File" in line:
            if not annot:
                pending = []
            annot = True
            for f, n in frame.findall(line):
                fr = (os.path.basename(f), int(n))
                if not pending or pending[-1] != fr:
                    pending.append(fr)
            continue
        annot = False
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S.*?);", line)
        if m:
            out.append((tuple(reversed(pending)), m.group(1)))
    return out


def main():
    src, lib = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else "probe_update_wavefrontILb0ELb0"
    rays = float(sys.argv[4]) if len(sys.argv) > 4 else 8388608.0
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    ch = chains(lib, want)
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[rows.index(hdr) + 1:]
    assert len(data) == len(ch), f"{len(data)} CSV rows vs {len(ch)} instructions: not the profiled build"
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    tot_i = tot_s = 0.0
    for r, (chain, _) in zip(data, ch):
        inst, thr, smp = float(r[ix["Instructions Executed"]] or 0), float(r[ix["Thread Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)
        a = agg[chain[:depth]]
        a[0] += 1; a[1] += inst; a[2] += thr; a[3] += smp
        tot_i += inst; tot_s += smp
    print(f"{want}: {len(ch)} SASS, {tot_i / rays:.1f} warp instructions per ray")
    print(f"{'SASS':>5s} {'warp-inst/ray':>13s} {'%':>6s} {'lanes':>6s} {'samples %':>9s}  line")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot_i < 0.002:
            continue
        print(f"{a[0]:5d} {a[1] / rays:13.2f} {100 * a[1] / tot_i:6.2f} {a[2] / a[1] if a[1] else 0:6.1f} {100 * a[3] / tot_s:9.2f}  " + "  <-  ".join(f"{f}:{n}" for f, n in key))


if __name__ == "__main__":
    main()
