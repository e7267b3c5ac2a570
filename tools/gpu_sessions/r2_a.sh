#!/bin/bash
# round 2, session A: first run of the lean K2 (pl_k2_lean): parity, A/B against the generic kernel, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
free -g >> gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lean" > gpurun_out/r2a_pytest_lean.log 2>&1; echo "pytest lean rc=$?"; tail -3 gpurun_out/r2a_pytest_lean.log
timeout 300 python tools/sweep.py --height 135 --images 2368,3552 --lanes 1 --bm 1 --lean 0,1 > gpurun_out/r2a_sweep.txt 2>&1; echo "sweep rc=$?"
timeout 200 python tools/sweep.py --height 135 --images 148,592,1184 --lanes 1 --bm 1 --lean 1 >> gpurun_out/r2a_sweep.txt 2>&1
cat gpurun_out/r2a_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/r2a_k2_lean python tools/sweep.py --height 24 --images 3552 --lanes 1 --bm 1 --lean 1 --reps 0 > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2a_ncu.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2a_pytest_all.log
