#!/bin/bash
# r2l: what a 1/8 share loses per ray against the full field: ncu of both (same build), and ownership units on one GPU
out=gpurun_out/r2l; mkdir -p $out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio
for w in 1 8; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:probe_update_wavefront -s 4 -c 1 --csv --log-file $out/share_$w.csv python profiles/diag_share_run.py field_32 $w 6 > /dev/null 2>&1
done
for u in 32 1024; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:probe_update_wavefront -s 4 -c 1 --csv --log-file $out/share_8_unit$u.csv python profiles/diag_share_run.py field_32 8 6 $u > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_update_wavefront -s 4 -c 1 -f -o $out/prof_share8 python profiles/diag_share_run.py field_32 8 6 > $out/ncu_share8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_update_wavefront -s 4 -c 1 -f -o $out/prof_share1 python profiles/diag_share_run.py field_32 1 6 > $out/ncu_share1.log 2>&1
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2l/share_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    print(f)
    for r in rows[1:]:
        print('   ', r[-3], r[-1], r[-2])
PY
