#!/usr/bin/env python
"""Static SASS profile of one kernel (no GPU): instructions per source line of the kernel body and, one level
down, per line of the inlined per-ray functions (nvdisasm -gi on the cubin extracted from the built library).
Tells where the instruction budget of each state of the probe-update state machine goes.

    python profiles/sass_by_line.py [kernel-substring] [depth]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dynamic-diffuse-global-illumination-minecraft_b200", "libddgi_b200.so")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "probe_update_wavefrontILb0ELb0"
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=d, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if f.startswith("ddgi_kernels.")][0]
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(d, cub)], capture_output=True, text=True).stdout
    in_fn = False
    chain = []
    counts = collections.Counter()
    total = 0
    frame = re.compile(r'File "([^"]+)", line (\d+)')
    pending = []
    annot = False
    for line in txt.splitlines():
        if line.startswith(".text."):
            in_fn = want in line
            continue
        if not in_fn:
            continue
        if "//## File" in line:
            # consecutive annotation lines spell one inline chain, innermost frame first
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
            if pending:
                chain = pending  # innermost first ... outermost last
            key = tuple(reversed(chain))[:depth]
            counts[key] += 1
            total += 1
    print(f"{want}: {total} SASS instructions")
    for key, c in sorted(counts.items(), key=lambda kv: kv[0]):
        print(f"{c:6d}  " + "  <-  ".join(f"{f}:{n}" for f, n in key))


if __name__ == "__main__":
    main()
