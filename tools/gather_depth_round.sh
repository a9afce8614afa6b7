#!/bin/bash
# N-GPU visit: the bench at N with the peer gather at several pipeline depths (+ the publish-only experiment and no gather at all).
N=${1:-2}; shift
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for v in "$@"; do
  i=$((i+1))
  timeout 300 $RUN --master-port $((29520+i)) bench.py --gpus $N --steps 200 --warmup 20 $v > "gpurun_out/bench_n${N}_${v// /_}.json" 2>> gpurun_out/bench_n$N.err
done
for f in gpurun_out/bench_n$N*.json; do python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(j['value']), round(j['ms_per_step']*1e3,2), round(j['e2e']['value']), j['run'].get('gather_timeouts'))
except Exception as e: print('$f', 'FAILED', e)
"; done; tail -5 gpurun_out/bench_n$N.err
