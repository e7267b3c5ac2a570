#!/bin/bash
# session X: bucket maxima with 32-bit relative keys (native ATOMS.MAX) - parity, sweep, bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_x.log
echo "== sweep 1184/2368 lanes 1,2 bm 0,1"
timeout 600 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1,2 --bm 1 > gpurun_out/sweep_x_a.log 2>&1; cut -c1-170 gpurun_out/sweep_x_a.log
timeout 600 python tools/sweep.py --height 135 --images 148,296,592 --lanes 8,4,2,1 --bm 1 > gpurun_out/sweep_x_b.log 2>&1; cut -c1-170 gpurun_out/sweep_x_b.log
echo "== profile build: per-filter busy"
