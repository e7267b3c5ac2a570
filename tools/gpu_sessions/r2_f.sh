#!/bin/bash
# round 2, session F: lean K2 after the symbol-0 bucket fix; tile / ring-depth variants
mkdir -p gpurun_out
for v in "" v_t8s4 v_t8s6; do
  lib=pngloss_b200/libpngloss_b200.so; [ -n "$v" ] && lib=pngloss_b200/lib$v.so
  echo "== variant ${v:-default t16s2}"
  PNGLOSS_B200_LIB=$PWD/$lib timeout 300 python tools/sweep.py --height 135 --images 148,2368,3552 --lanes 1 --bm 1 --lean 1 2>&1 | cut -c1-200
done > gpurun_out/r2f_sweep.txt 2>&1
timeout 200 python tools/sweep.py --height 135 --images 148 --lanes 1,8 --bm 1 --lean 0 >> gpurun_out/r2f_sweep.txt 2>&1
cat gpurun_out/r2f_sweep.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lean" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/r2f_k2_lean python tools/sweep.py --height 24 --images 3552 --lanes 1 --bm 1 --lean 1 --reps 0 > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
