#!/bin/bash
mkdir -p gpurun_out
echo "== per-filter busy cycles (profile build)"
PNGLOSS_B200_LIB=$PWD/pngloss_b200/exp_profile.so timeout 600 python tools/sweep.py --height 135 --images 148,1184 --lanes 8,2,1 --reps 0 --profile > gpurun_out/profile_o.log 2>&1; grep busy gpurun_out/profile_o.log
echo "== ncu K2 lanes 1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_o_k2_l1 python tools/sweep.py --height 32 --images 1184 --lanes 1 --reps 0 > gpurun_out/ncu_o_k2_l1.log 2>&1; echo "rc=$?"
