#!/bin/bash
# 8-GPU box (gpurun --gpus 8): the bench line at 8 and 4 GPUs (both pipelines, --verify, NCCL line), then the reference arm at 8
tag=${1:-scale8}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.txt; nproc > $out/nproc.txt
for n in ${NS:-8 4}; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --verify 2> $out/bench_n$n.err | grep "^{" > $out/bench_n$n.json
done
[ "${SKIP_REF:-0}" = "1" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --impl reference --steps 5 --warmup 3 2> $out/bench_reference_n8.err | grep "^{" > $out/bench_reference_n8.json
for f in $out/bench_n*.json; do echo $f; python - "$f" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value %.4g  ms %.3f kernel_ms %.3f e2e %.4g fps %.1f verify %s nccl %s pipelines %s e2e_by %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['value'], j['fps']['value'], j.get('verify'), (j.get('exchange_nccl') or {}).get('value'), {k: (v or {}).get('value') if isinstance(v, dict) else v for k, v in (j.get('pipelines') or {}).items()}, j['e2e'].get('by_frames_in_flight')))
PY
done
for f in $out/*.err; do tail -n 2 $f | cut -c1-300; done
