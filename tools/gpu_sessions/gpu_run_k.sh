#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_k.log
echo "== sweep"; timeout 900 python tools/sweep.py --height 135 --images 592,1184,1776 --lanes 4,2,1 > gpurun_out/sweep_k.log 2>&1; cut -c1-150 gpurun_out/sweep_k.log
echo "== bench full trace"; PNGLOSS_BENCH_TRACE=1 timeout 1500 python bench.py --no-cpu > gpurun_out/bench_k.log 2>&1; grep trace gpurun_out/bench_k.log; tail -1 gpurun_out/bench_k.log | cut -c1-1300
