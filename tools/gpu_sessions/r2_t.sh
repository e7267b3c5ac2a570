#!/bin/bash
# round 2, session T: quick A/B of the latency kernel (one image, 148 and 296 images) after a change
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "solo" 2>&1 | tail -2
timeout 300 python tools/sweep.py --width 3840 --height 135 --images 1,148,296 --lanes 0 --solo -1 --reps 2 2>&1 | cut -c1-260 > gpurun_out/r2t_sweep.txt
cat gpurun_out/r2t_sweep.txt
