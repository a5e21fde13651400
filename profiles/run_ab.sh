#!/bin/bash
# One GPU call of a tuning pass (run under gpurun from the repo root): GPU tests, then A/B of the builds under profiles/ab
# against the in-tree build on the same box, then one ncu --set full capture of the probe-update kernel.
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
  DDGI_LIB=$lib timeout 300 python profiles/ab_kernel.py ${AB_WORKLOADS:-field_32,cave_128,cave_64} ${AB_VARIANTS:-1,2} ${AB_MM:-16} >> $out/ab.txt 2>&1
done
DDGI_LIB= timeout 300 python profiles/ab_kernel.py field_32 2 12,14,18,20 >> $out/ab.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_update_wavefront -s 6 -c 1 -f -o $out/prof_wf \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_full.log 2>&1
tail -5 $out/pytest_gpu.log; cat $out/ab.txt
