#!/bin/bash
# round 2, session K: A/B of the generic K2 with the all-active specialisation + direct bucket update against the
# previous build, same box
mkdir -p gpurun_out
for lib in libprev_generic.so libpngloss_b200.so; do
  echo "== $lib"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/$lib timeout 300 python tools/sweep.py --height 135 --images 148,2368 --lanes 1 --bm 1 --lean 0 --reps 2 2>&1 | cut -c1-230
done > gpurun_out/r2k_ab.txt 2>&1
cat gpurun_out/r2k_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
