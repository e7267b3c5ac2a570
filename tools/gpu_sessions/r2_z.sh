#!/bin/bash
# round 2, session Z: the latency kernel at strengths below 15 (tables with up to 259 buckets): parity, single-image time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "solo or golden_via_dropin" 2>&1 | tail -3
for s in 0 5 10 20; do
  timeout 300 python tools/sweep.py --width 3840 --height 135 --images 1,296 --lanes 0 --solo -1 --strength $s --reps 1
done 2>&1 | cut -c1-250 > gpurun_out/r2z_sweep.txt
cat gpurun_out/r2z_sweep.txt
timeout 400 python bench.py --config 3 --strength 0 --steps 2 --warmup 1 > gpurun_out/r2z_config3_s0.json 2> gpurun_out/r2z_config3_s0.err; echo "config 3 s0 rc=$?"; cut -c1-160 gpurun_out/r2z_config3_s0.json
