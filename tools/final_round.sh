#!/bin/bash
# The record of a round on one GPU: whole -m gpu suite, smoke, the bench (both arms, every config), launch list + full ncu of the scan.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k20.json 2>> gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
python bench.py --head-dtype f16 --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_f16.json 2>> gpurun_out/bench.err
for c in v5s_b1 retina_b16 rcnn_b8; do
  python bench.py --config $c --steps 200 --warmup 20 > gpurun_out/bench_$c.json 2>> gpurun_out/bench.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:yolo_v8_scan -s 8 -c 1 -o gpurun_out/scan -f python tools/ncu_target.py > gpurun_out/ncu_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:letterbox_unit -s 8 -c 1 -o gpurun_out/letterbox -f python tools/ncu_target.py > gpurun_out/ncu_lb.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/other_kernels.csv python bench.py --config rcnn_b8 --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2
for f in gpurun_out/bench.json gpurun_out/bench_k20.json gpurun_out/bench_reference.json gpurun_out/bench_f16.json gpurun_out/bench_v5s_b1.json gpurun_out/bench_retina_b16.json gpurun_out/bench_rcnn_b8.json; do python -c "
import json
try:
    j=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(j['value']), round(j['ms_per_step']*1e3,2), round(j['e2e']['value']), {k: round(v, 1) for k, v in (j.get('kernels_us') or {}).items()}, (j.get('roofline') or {}).get('frac'), (j.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('$f', 'FAILED', e)
"; done; tail -5 gpurun_out/bench.err
