#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, scan-variant sweep, ncu launch list + full capture of the scan kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== smoke" ; timeout 180 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 300 --warmup 20 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
echo "== bench --impl reference" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_reference.json
tail -5 $OUT/${TAG}_bench.err
echo "== sweep" ; timeout 600 python tools/sweep_scan.py 2>&1 | tee $OUT/${TAG}_sweep.log
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "== ncu full (scan, nms, letterbox)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"yolo_v8_scan|nms_kernel|letterbox" -s 9 -c 6 -f -o $OUT/${TAG}_full \
    python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
