#!/bin/bash
mkdir -p gpurun_out
for lib in exp_v0 exp_v1 libpngloss_b200; do
  echo "== $lib"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$lib.so timeout 600 python tools/sweep.py --height 135 --images 592 --lanes 8,2 > gpurun_out/sweep_s_${lib}_a.log 2>&1; cut -c1-110 gpurun_out/sweep_s_${lib}_a.log
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$lib.so timeout 600 python tools/sweep.py --height 135 --images 1184 --lanes 2,1 > gpurun_out/sweep_s_${lib}_b.log 2>&1; cut -c1-110 gpurun_out/sweep_s_${lib}_b.log
done
