#!/bin/bash
# session AE: row-commit variants (c0 scalar, c1 16-byte vectors, c2 = c1 not inlined); timeline of the job API
mkdir -p gpurun_out
for v in exp_c0 exp_c1 exp_c2; do
  echo "== $v"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$v.so timeout 300 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1 --bm 1 2>&1 | cut -c1-150
done
PNGLOSS_B200_LIB=$PWD/pngloss_b200/exp_c1.so timeout 300 python tools/sweep.py --height 135 --images 148,592 --lanes 8,2 --bm 0 2>&1 | cut -c1-150
echo "== e2e trace"
timeout 600 python tools/e2e_trace.py 2>&1 | tee gpurun_out/e2e_trace_ae.log
