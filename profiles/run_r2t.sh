#!/bin/bash
# r2t: the two-slot kernel (variants 3 / 4): GPU parity tests, then A/B against variant 2
out=gpurun_out/${1:-r2t}; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
DDGI_LIB=$lib timeout 300 python profiles/ab_kernel.py field_32,cave_128,cave_64,sweep_1024 2,4 16 >> $out/ab.txt 2>&1
done
DDGI_LIB= timeout 300 python profiles/ab_kernel.py field_32 4 8,12,20,24 >> $out/ab.txt 2>&1
cat $out/ab.txt
