#!/bin/bash
# r2o: how much of field_32 is uniform rays (probes inside rock), and what such a ray costs on its own (all-solid box)
out=gpurun_out/r2o; mkdir -p $out
timeout 300 python profiles/diag_rock.py field_32 > $out/rock.txt 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none -k regex:probe_update_wavefront -s 4 -c 1 --csv --log-file $out/solid.csv python profiles/diag_rock.py field_32 solid 6 >> $out/rock.txt 2>&1
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2o/solid_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    print(f, '; '.join(f"{r[-3].split('.')[0].replace('smsp__','').replace('gpu__','')} {r[-1]}" for r in rows[1:]))
PY
cat $out/rock.txt
