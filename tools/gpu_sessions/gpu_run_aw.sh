#!/bin/bash
# session AW: ncu --set full of the final K2 (2368 x 3840x24)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_aw_k2_final python tools/sweep.py --height 24 --images 2368 --lanes 1 --bm 1 --reps 0 > gpurun_out/ncu_aw_k2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_aw_k2.log
