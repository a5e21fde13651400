#!/bin/bash
# One 1-GPU measurement pass of a round (run under gpurun from the repo root):
#   GPU tests, smoke, the bench line (both arms), the other BASELINE configs, the ncu launch list
#   and one --set full capture of the probe-update kernel.  Outputs under gpurun_out/$1/.
tag=${1:-round}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 300 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
for w in cave_64 cave_128 sweep_64 sweep_128 sweep_256 sweep_512 sweep_1024; do
  timeout 120 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out/other_configs.jsonl
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_update_wavefront -s 6 -c 1 -f -o $out/prof_wf \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/ncu_full.log 2>&1
tail -3 $out/pytest_gpu.log; tail -1 $out/smoke.log; cut -c1-300 $out/bench_n1.json; ls -la $out
