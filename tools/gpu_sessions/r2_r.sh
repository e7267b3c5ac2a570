#!/bin/bash
# round 2, session R: the latency kernel with several CTAs per SM (chain warps rotated over the sub-partitions)
# against the generic kernel's lane mappings at the same batch sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "solo" 2>&1 | tail -3
{
timeout 600 python tools/sweep.py --width 3840 --height 135 --images 148,296,444,592,740 --lanes 8 --solo 1 --reps 1
timeout 600 python tools/sweep.py --width 3840 --height 135 --images 296,592,740 --lanes 0 --solo 0 --reps 1
timeout 600 python tools/sweep.py --width 1920 --height 270 --images 128,1024 --lanes 8 --solo 1 --reps 1
timeout 600 python tools/sweep.py --width 1920 --height 270 --images 128,1024 --lanes 0 --solo 0 --reps 1
} 2>&1 | cut -c1-260 > gpurun_out/r2r_sweep.txt
cat gpurun_out/r2r_sweep.txt
