#!/bin/bash
# session AA: which of the last changes cost 10 % of K2?  v0 = previous commit, v2 = plain barrier,
# v3 = no oprev row, v4 = both
mkdir -p gpurun_out
for v in exp_v0 libpngloss_b200 exp_v2 exp_v3 exp_v4; do
  echo "== $v"
  PNGLOSS_B200_NO_SPARE_WARP=1 PNGLOSS_B200_LIB=$PWD/pngloss_b200/$v.so timeout 300 python tools/sweep.py --height 135 --images 1184 --lanes 1 --bm 1 2>&1 | cut -c1-150
done
