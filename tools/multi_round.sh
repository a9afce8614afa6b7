#!/bin/bash
# N-GPU visit (gpurun --gpus N): peer-gather correctness, then the bench at N with the fused gather and with NCCL for comparison.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_PEER_CHECK" ]; then timeout 300 $RUN --master-port 29511 tools/peer_gather_check.py > gpurun_out/peer_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/peer_check_n$N.log; fi
timeout 600 $RUN --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
for v in "--no-gather" "--fused-gather" "--nccl-gather"; do
  timeout 400 $RUN --master-port 29514 bench.py --gpus $N --steps 200 --warmup 20 $v > "gpurun_out/bench_n${N}_${v// /_}.json" 2>> gpurun_out/bench_n$N.err
done
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n1.json 2>> gpurun_out/bench_n$N.err
tail -3 gpurun_out/peer_check_n$N.log; for f in gpurun_out/bench_n1.json gpurun_out/bench_n$N*.json; do python -c "
import json,sys
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(j['value']), round(j['ms_per_step']*1e3,2), round(j['e2e']['value']), j['run']['parallelism'][:60], j['run'].get('gather_timeouts'), {k: round(v, 1) for k, v in j['kernels_us'].items() if 'nms' in k})
except Exception as e: print('$f', 'FAILED', e)
"; done; tail -5 gpurun_out/bench_n$N.err
