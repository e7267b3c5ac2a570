#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r.log
echo "== sweep"; timeout 600 python tools/sweep.py --height 135 --images 148,592,1184 --lanes 8,4,2,1 > gpurun_out/sweep_r.log 2>&1; cut -c1-150 gpurun_out/sweep_r.log
echo "== ncu K2 lanes 1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_r_k2_l1 python tools/sweep.py --height 32 --images 1184 --lanes 1 --reps 0 > gpurun_out/ncu_r_k2_l1.log 2>&1; echo "rc=$?"
