#!/bin/bash
mkdir -p gpurun_out
echo "== ncu K2 lanes 8"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_k2_l8 python tools/sweep.py --height 32 --images 444 --lanes 8 --reps 0 > gpurun_out/ncu_k2_l8.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_k2_l8.log
echo "== ncu K2 lanes 2"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_k2_l2 python tools/sweep.py --height 32 --images 1776 --lanes 2 --reps 0 > gpurun_out/ncu_k2_l2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_k2_l2.log
echo "== ncu K1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k1 -c 1 -f -o gpurun_out/prof_k1 python tools/sweep.py --height 2160 --images 32 --lanes 8 --reps 0 > gpurun_out/ncu_k1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_k1.log
ls -la gpurun_out/*.ncu-rep
