#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_c.log
for lib in libpngloss_b200 exp_mb3 exp_mb5 exp_mb6; do
  echo "== sweep $lib"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$lib.so timeout 600 python tools/sweep.py --height 135 --images 592,888,1184 --lanes 8,4,2 > gpurun_out/sweep_c_$lib.log 2>&1
  cat gpurun_out/sweep_c_$lib.log | cut -c1-140
done
echo "== ncu K2 lanes 8"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_c_k2_l8 python tools/sweep.py --height 32 --images 592 --lanes 8 --reps 0 > gpurun_out/ncu_c_k2_l8.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_c_k2_l8.log
