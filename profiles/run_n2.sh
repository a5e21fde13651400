#!/bin/bash
# 2-GPU check of the bench line (run under gpurun --gpus 2): fused exchange with two frames in flight, one frame at a time, and the NCCL line
tag=${1:-n2}; out=gpurun_out/$tag; mkdir -p $out
for fl in 0; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$fl bench.py --gpus 2 --verify --no-ncu 2> $out/bench_n2_fl$fl.err | grep "^{" > $out/bench_n2_fl$fl.json
  tail -3 $out/bench_n2_fl$fl.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "shard or peer or exchange or multi or nccl or flight" > $out/pytest_multi.log 2>&1; tail -3 $out/pytest_multi.log
for f in $out/bench_n2_fl*.json; do echo $f; python - "$f" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value %.4g  ms %.3f kernel_ms %.3f e2e %.4g fps %.1f verify %s nccl %s pipelines %s e2e_by %s' % (j['value'], j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['value'], j['fps']['value'], j.get('verify'), (j.get('exchange_nccl') or {}).get('value'), {k: (v or {}).get('value') if isinstance(v, dict) else v for k, v in (j.get('pipelines') or {}).items()}, j['e2e'].get('by_frames_in_flight')))
PY
done
