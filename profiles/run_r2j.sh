#!/bin/bash
# r2j: GPU tests on the MSB occupancy layout, A/B of the march-loop trims, frames-in-flight on a 1/N share, ncu of the TMA kernel
out=gpurun_out/r2j; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
AB_WORKLOADS=field_32,cave_128,cave_64,sweep_1024
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
  DDGI_LIB=$lib timeout 300 python profiles/ab_kernel.py $AB_WORKLOADS 2 16 >> $out/ab.txt 2>&1
done
timeout 300 python profiles/diag_inflight.py field_32 1,2,4,8 40 > $out/inflight.txt 2>&1
(cd profiles/experiments/tma_staging && timeout 300 ncu --set full --clock-control none --import-source on -k regex:primary_march -c 4 -f -o ../../../$out/prof_tma ./tma_march 16 1 > ../../../$out/ncu_tma.log 2>&1)
tail -3 $out/pytest_gpu.log; cat $out/ab.txt $out/inflight.txt
