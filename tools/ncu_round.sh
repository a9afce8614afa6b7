#!/bin/bash
# ncu pass (one GPU): full-set captures of the scan and letterbox kernels (warm launches, rotating inputs) + the launch list of the bench.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:yolo_v8_scan -s 8 -c 2 -o gpurun_out/r02_scan -f python tools/ncu_target.py > gpurun_out/ncu_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:letterbox_unit -s 8 -c 2 -o gpurun_out/r02_letterbox -f python tools/ncu_target.py > gpurun_out/ncu_lb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_kernel -s 8 -c 1 -o gpurun_out/r02_nms -f python tools/ncu_target.py > gpurun_out/ncu_nms.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/sweep_scan.py > gpurun_out/sweep.log 2>&1
python tools/reg_probe.py > gpurun_out/reg_probe.log 2>&1
tail -3 gpurun_out/ncu_scan.log; tail -25 gpurun_out/sweep.log; tail -40 gpurun_out/reg_probe.log
