#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_d.log
echo "== sweep"; timeout 900 python tools/sweep.py --height 135 --images 148,296,592,1184,1776 --lanes 8,4,2,1 > gpurun_out/sweep_d.log 2>&1; cut -c1-150 gpurun_out/sweep_d.log
echo "== bench 1184 lanes 2"; timeout 1200 python bench.py --images 1184 --lanes 2 --steps 2 --warmup 3 > gpurun_out/bench_d.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_d.log
