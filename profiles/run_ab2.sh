#!/bin/bash
# A/B of the builds under profiles/ab against the in-tree build (run under gpurun): bash profiles/run_ab2.sh <tag> [workloads] [variants] [march_mins]
tag=${1:-ab}; out=gpurun_out/$tag; mkdir -p $out
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
  DDGI_LIB=$lib timeout 300 python profiles/ab_kernel.py ${2:-field_32,cave_128} ${3:-2} ${4:-16} >> $out/ab.txt 2>&1
done
cat $out/ab.txt
