#!/bin/bash
# session AC: bisect the 9 % K2 slow-down: v0 = library of the previous commit, v5 = launch bounds 160,
# v6 = launch bounds 160 + no oprev + plain barrier
mkdir -p gpurun_out
echo "== exp_v0 (previous commit)"
PNGLOSS_B200_OLD_LIB=1 PNGLOSS_B200_LIB=$PWD/pngloss_b200/exp_v0.so timeout 300 python tools/sweep.py --height 135 --images 1184 --lanes 1 --bm 1 2>&1 | cut -c1-150
for v in exp_v5 exp_v6 exp_v4; do
  echo "== $v"
  PNGLOSS_B200_NO_SPARE_WARP=1 PNGLOSS_B200_LIB=$PWD/pngloss_b200/$v.so timeout 300 python tools/sweep.py --height 135 --images 1184 --lanes 1 --bm 1 2>&1 | cut -c1-150
done
