#!/bin/bash
# session AJ: bucket-maxima fast path without the accumulator merge; A/B against the committed library (exp_prev)
mkdir -p gpurun_out
for v in exp_prev libpngloss_b200; do
  echo "== $v"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$v.so timeout 300 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1,2 --bm 1 2>&1 | cut -c1-150
done
