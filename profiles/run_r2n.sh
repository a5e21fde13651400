#!/bin/bash
# r2n: lockstep batches - (lockstep) uniform tail of the schedule on / off: GPU tests, kernel A/B, 1/8 share, ncu counters
out=gpurun_out/r2n; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
for g in 1 0; do
  DDGI_GATHER=$g timeout 300 python profiles/ab_kernel.py field_32,cave_128,cave_64,sweep_1024 2 16 >> $out/ab.txt 2>&1
  DDGI_GATHER=$g timeout 300 python profiles/diag_inflight.py field_32 1,8 40 >> $out/inflight.txt 2>&1
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
for w in 1 8; do for g in 1 0; do
  DDGI_GATHER=$g timeout 300 ncu --metrics $M --clock-control none -k regex:probe_update_wavefront -s 4 -c 1 --csv --log-file $out/share_${w}_g$g.csv python profiles/diag_share_run.py field_32 $w 6 > /dev/null 2>&1
done; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_update_wavefront -s 4 -c 1 -f -o $out/prof_wf python profiles/diag_share_run.py field_32 1 6 > $out/ncu_full.log 2>&1
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2n/share_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    print(f, '; '.join(f"{r[-3].split('.')[0].replace('smsp__','').replace('gpu__','')} {r[-1]}" for r in rows[1:]))
PY
tail -3 $out/pytest_gpu.log; cat $out/ab.txt $out/inflight.txt
