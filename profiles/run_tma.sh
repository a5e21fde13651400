#!/bin/bash
# The TMA staging experiment (run under gpurun): timing + ncu stall picture of both kernels.
root=$(pwd); out=$root/gpurun_out/${1:-tma}; mkdir -p $out
cd profiles/experiments/tma_staging
./tma_march 32 15 > $out/tma_march.txt 2>&1
./tma_march 16 15 >> $out/tma_march.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:primary_march -c 2 -f -o $out/prof_tma ./tma_march 16 1 > $out/ncu_tma.log 2>&1
cd $root
cat $out/tma_march.txt
