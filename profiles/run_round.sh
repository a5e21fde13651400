#!/bin/bash
# One measurement pass of a round (run under gpurun from the repo root; N GPUs = $2, default 1):
#   GPU tests, smoke, the bench line (both arms), N-GPU bench lines with --verify, the other BASELINE configs,
#   the ncu launch list and one --set full capture of the probe-update and the pixel kernels.  Outputs under gpurun_out/$1/.
tag=${1:-round}
ngpu=${2:-1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt
nproc > $out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 400 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
for n in 2 4 8; do
  if [ $n -le $ngpu ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --verify > $out/bench_n$n.json 2> $out/bench_n$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --impl reference --steps 5 > $out/bench_reference_n$n.json 2> $out/bench_reference_n$n.err
  fi
done
if [ "${SKIP_OTHER:-0}" != "1" ]; then
for w in cave_64 cave_128 sweep_64 sweep_128 sweep_256 sweep_512 sweep_1024; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-ncu 2>/dev/null | tail -1 >> $out/other_configs.jsonl
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > $out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_update_wavefront -s 6 -c 1 -f -o $out/prof_wf \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > $out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_frame_kernel -s 2 -c 1 -f -o $out/prof_px \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > $out/ncu_full_px.log 2>&1
fi
tail -3 $out/pytest_gpu.log; tail -1 $out/smoke.log; for f in $out/bench_*.json; do echo $f; cut -c1-400 $f; done; ls -la $out
