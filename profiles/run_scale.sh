#!/bin/bash
# Scaling lines on one box (run under gpurun --gpus N): bench.py at 1 and at every power of two up to $2 GPUs, with --verify.
tag=${1:-scale}; ngpu=${2:-8}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.txt
timeout 300 python bench.py --no-cpu-baseline --no-ncu > $out/bench_n1.json 2> $out/bench_n1.err
for n in 2 4 8; do
  if [ $n -le $ngpu ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --verify 2> $out/bench_n$n.err | grep "^{" > $out/bench_n$n.json
  fi
done
for f in $out/bench_n*.json; do echo $f; python - "$f" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value %.4g  ms %.3f kernel_ms %.3f e2e %.4g fps %.1f verify %s nccl %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['value'], j['fps']['value'], j.get('verify'), (j.get('exchange_nccl') or {}).get('value')))
PY
done
