#!/bin/bash
# session AT: A/B of register budget (min blocks 3 -> 124 registers) and of the fall-back scan's unroll factor
mkdir -p gpurun_out
for v in libpngloss_b200 exp_mb3 exp_unr4 exp_unr6; do
  echo "== $v"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$v.so timeout 300 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1 --bm 1 2>&1 | cut -c1-150
done
