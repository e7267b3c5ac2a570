#!/bin/bash
# round 2, session O: first GPU run of the latency kernel (pl_k2_solo): parity tests, then single-image and
# one-image-per-SM timings against the generic kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "solo" 2>&1 | tail -5
for h in 135; do
  timeout 300 python tools/sweep.py --width 3840 --height $h --images 1,148 --lanes 8 --solo 0,1,2 --reps 2
done > gpurun_out/r2o_sweep.txt 2>&1
timeout 120 python tools/sweep.py --width 512 --height 512 --images 1 --lanes 8 --solo 0,1,2 --reps 2 >> gpurun_out/r2o_sweep.txt 2>&1
timeout 120 python tools/sweep.py --width 180 --height 215 --images 1 --lanes 8 --solo 0,1,2 --reps 2 >> gpurun_out/r2o_sweep.txt 2>&1
cat gpurun_out/r2o_sweep.txt
