#!/bin/bash
# r2k: block size of the persistent kernel (128 / 64 / 32 threads) on the full field and on a 1/8 share, one and two frames in flight
out=gpurun_out/r2k; mkdir -p $out
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
  DDGI_LIB=$lib timeout 300 python profiles/ab_kernel.py field_32,cave_128,cave_64 2 16 >> $out/ab.txt 2>&1
  DDGI_LIB=$lib timeout 300 python profiles/diag_inflight.py field_32 1,4,8 40 >> $out/inflight.txt 2>&1
done
cat $out/ab.txt $out/inflight.txt
