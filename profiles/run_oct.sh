#!/bin/bash
# TMA-staged octahedral blend: parity tests of the mode, A/B against the plain-load build, ncu of the blend kernel
out=gpurun_out/${1:-oct}; mkdir -p $out
timeout 600 python -m pytest tests/test_octahedral.py tests/test_modes.py -m gpu -q -x > $out/pytest_oct.log 2>&1; tail -3 $out/pytest_oct.log
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
  for w in field_32 cave_128; do DDGI_LIB=$lib timeout 300 python profiles/ab_oct.py $w 8 15 >> $out/ab_oct.txt 2>&1; done
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
for lib in "" $(ls profiles/ab/*.so 2>/dev/null); do
  DDGI_LIB=$lib timeout 300 ncu --metrics $M --clock-control none -k regex:probe_blend_octahedral -s 3 -c 1 --csv --log-file $out/blend_$(basename ${lib:-default} .so).csv python profiles/ab_oct.py field_32 8 3 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/*/blend_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    print(f, '; '.join(f"{r[-3].split('.')[0]} {r[-1]}" for r in rows[1:]))
PY
cat $out/ab_oct.txt
