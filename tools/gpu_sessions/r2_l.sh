#!/bin/bash
# round 2, session L: (1) does the bucket table pay for wide lane groups at high strengths (single image and 148)?
# (2) the default bench line exactly as the driver runs it
mkdir -p gpurun_out
for s in 40 85; do
  timeout 300 python tools/sweep.py --height 135 --images 1,148 --lanes 8,4,1 --bm 0,1 --lean 0 --strength $s --reps 1 2>&1 | cut -c1-200
done > gpurun_out/r2l_strength_bm.txt 2>&1
cat gpurun_out/r2l_strength_bm.txt
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_default.json 2> gpurun_out/r2l_default.err ) 2>&1 | tail -3; echo "default rc=$?"; cat gpurun_out/r2l_default.json | cut -c1-1500; tail -3 gpurun_out/r2l_default.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_reference.json 2> gpurun_out/r2l_reference.err ) 2>&1 | tail -3; cat gpurun_out/r2l_reference.json | cut -c1-700
