#!/bin/bash
# round 2, session Q: A/B of latency-kernel build variants (one vote / two votes, third table entry uniform / always)
mkdir -p gpurun_out
for v in 00 10 01 11; do
  echo "== variant onevote/three_uniform = $v"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/libsolo$v.so timeout 300 python tools/sweep.py --width 3840 --height 135 --images 1 --lanes 8 --solo 1,2 --reps 2 | cut -c1-150
done > gpurun_out/r2q_variants.txt 2>&1
cat gpurun_out/r2q_variants.txt
