#!/bin/bash
# compute-sanitizer over every kernel and mode (profiles/sanitize_small.py): memcheck, racecheck, synccheck
out=gpurun_out/${1:-san2}; mkdir -p $out
python profiles/sanitize_small.py > $out/plain.txt 2>&1; echo "plain rc=$?" >> $out/plain.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python profiles/sanitize_small.py > $out/$tool.txt 2>&1; echo "$tool rc=$?" >> $out/$tool.txt
done
for f in plain memcheck racecheck synccheck; do echo "== $f"; tail -4 $out/$f.txt; done
