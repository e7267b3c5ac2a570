#!/bin/bash
# round 2, session Y: K4 groups per thread with the conflict-free staging
mkdir -p gpurun_out
for v in k4g1 "" k4g3 k4g4; do
  lib=pngloss_b200/libpngloss_b200.so; [ -n "$v" ] && lib=pngloss_b200/lib$v.so
  echo "== ${v:-default (2 groups)}"
  PNGLOSS_B200_LIB=$PWD/$lib timeout 300 python tools/k4_bench.py --images 1184 --height 540
done > gpurun_out/r2y_k4.txt 2>&1
cat gpurun_out/r2y_k4.txt
