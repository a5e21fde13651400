#!/usr/bin/env python
"""Splits an ncu source-page CSV of one kernel by OUTERMOST source line (for the wavefront
kernel: by state of the per-lane state machine) and by innermost function.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libddgi_b200.so; nvdisasm -gi -c ddgi_kernels.sm_100a.cubin > k.dis
    python profiles/prof_by_state.py src.csv k.dis probe_update_wavefront [n_rays]

`nvdisasm -gi` prints before each SASS instruction its inline chain, innermost line
first and the kernel's own line last; the n-th instruction of the kernel is the n-th CSV
row.  The kernel line is mapped to a label by the LABELS table (source text matching).
"""
import csv
import re
import sys
from collections import defaultdict

STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_not_selected", "stall_selected",
          "stall_branch_resolving", "stall_no_inst", "stall_lg", "stall_mio", "stall_dispatch", "stall_barrier"]


def label_of(kernel_src, line):
    """Label of a kernel source line: the nearest preceding `// ----` comment or state call."""
    text = kernel_src[line - 1] if 0 < line <= len(kernel_src) else ""
    for pat, name in (("wf_step(", "MARCH"), ("wf_begin_query", "QUERY"), ("wf_resolve_hit", "HIT"), ("wf_end_march", "MARCH"), ("wf_resolve_bounce", "BOUNCE_HIT"),
                      ("wf_resolve_feeler", "FEELER_HIT"), ("wf_scatter", "SCATTER"), ("wf_step_literal", "MARCH_SLOW"),
                      ("store_texel", "FETCH"), ("fetch_ray", "FETCH"), ("wf_init", "FETCH"),
                      ("wf_query_block", "QUERY")):
        if pat in text:
            return name
    return None


def main():
    src_csv, dis, kernel = sys.argv[1:4]
    n_rays = float(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    chains = []
    group = []
    inside = False
    fresh = True
    for line in open(dis):
        if line.startswith(".text."):
            inside = kernel in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            if fresh:
                group = []
                fresh = False
            group.append((m.group(1), int(m.group(2))))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", line):
            chains.append(list(group))
            fresh = True
    if len(chains) != len(data):
        print(f"warning: {len(chains)} instructions in the disassembly vs {len(data)} CSV rows", file=sys.stderr)
    kfile = None
    for ch in chains:
        if ch:
            kfile = ch[-1][0]
            break
    ksrc = open(kfile).read().split("\n") if kfile else []

    def col(r, name):
        i = ix.get(name)
        if i is None or r[i] == "":
            return 0.0
        return float(r[i])

    agg = defaultdict(lambda: defaultdict(float))
    tot = defaultdict(float)
    for ch, r in zip(chains, data):
        outer = ch[-1] if ch else ("?", 0)
        lab = label_of(ksrc, outer[1]) if outer[0] == kfile else None
        if lab is None:
            lab = "scheduler"  # ballots, match/redux, loop control of the kernel body
        a = agg[lab]
        a["sass"] += 1
        for k, name in (("warp", "Instructions Executed"), ("thread", "Thread Instructions Executed"),
                        ("samples", "# Samples")):
            v = col(r, name)
            a[k] += v
            tot[k] += v
        for s in STALLS:
            v = col(r, s)
            a[s] += v
            tot[s] += v
    print(f"kernel {kernel}: {len(data)} SASS instructions, {tot['warp']:.4g} warp instructions, "
          f"{tot['thread'] / max(tot['warp'], 1):.2f} active threads / instruction"
          + (f", {tot['thread'] / n_rays:.0f} thread instructions / ray, {tot['warp'] / n_rays:.0f} warp instructions / ray" if n_rays else ""))
    print(f"{'state':12s} {'SASS':>5s} {'warp-inst %':>11s} {'thr/inst':>8s} {'samples %':>9s}   top stall reasons (share of the state's samples)")
    for lab, a in sorted(agg.items(), key=lambda kv: -kv[1]["warp"]):
        st = sorted(((a[s], s) for s in STALLS), reverse=True)[:4]
        ssum = sum(a[s] for s in STALLS) or 1.0
        stxt = ", ".join(f"{s[6:]} {100 * v / ssum:.0f}%" for v, s in st if v > 0)
        print(f"{lab:12s} {int(a['sass']):5d} {100 * a['warp'] / tot['warp']:11.1f} {a['thread'] / max(a['warp'], 1):8.1f} "
              f"{100 * a['samples'] / max(tot['samples'], 1):9.1f}   {stxt}")
    ssum = sum(tot[s] for s in STALLS) or 1.0
    print("all states: " + ", ".join(f"{s[6:]} {100 * tot[s] / ssum:.0f}%" for s in STALLS if tot[s] / ssum > 0.01))


if __name__ == "__main__":
    main()
