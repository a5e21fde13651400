#!/bin/bash
# warp instructions of ONE rank's share for world = 1 .. 32: is the inflation per ray a fixed cost per launch?
out=gpurun_out/${1:-share_scan}; mkdir -p $out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in 1 2 4 8 16 32; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:probe_update_wavefront -s 4 -c 1 --csv --log-file $out/share_$w.csv python profiles/diag_share_run.py field_32 $w 6 5 > /dev/null 2>&1
done
python - <<PY
import csv
for w in (1,2,4,8,16,32):
    rows=[r for r in csv.reader(open("$out/share_%d.csv" % w)) if len(r)>5]
    m={r[-3]:float(r[-1].replace(",","")) for r in rows[1:]}
    n=8388608//w
    print("1/%-2d share: %8d rays  %.3f ms  warp-inst %.1f M = %.1f per ray  thread-inst per ray %.0f  lanes %.2f  issue %.1f %%" % (w, n, m["gpu__time_duration.sum"]/1e6, m["smsp__inst_executed.sum"]/1e6, m["smsp__inst_executed.sum"]/n, m["smsp__thread_inst_executed.sum"]/n, m["smsp__thread_inst_executed.sum"]/m["smsp__inst_executed.sum"], m["smsp__issue_active.avg.pct_of_peak_sustained_active"]))
PY
