#!/bin/bash
# One GPU visit: the whole -m gpu suite, the letterbox debug cases, then the timing probes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python tools/lb_debug.py > gpurun_out/lb_debug.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/lb_probe.py > gpurun_out/lb_probe.log 2>&1
python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 300 --warmup 20 --no-overlap --no-cpu-baseline > gpurun_out/bench_nooverlap.json 2>> gpurun_out/bench.err
tail -5 gpurun_out/lb_debug.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/lb_probe.log; cat gpurun_out/bench.json | cut -c1-600
