#!/bin/bash
# session AN: K4 filter kernel at 8 resident CTAs per SM (32 registers) vs 6
mkdir -p gpurun_out
for v in libpngloss_b200 exp_k4b8; do
  echo "== $v"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$v.so timeout 600 python tools/k4_bench.py --images 296 --height 2160 2>&1 | tail -1
done
