#!/bin/bash
# session Y: per-filter busy cycles (profile build) and the default bench with the 32-bit-key bucket maxima
mkdir -p gpurun_out
PNGLOSS_B200_LIB=$PWD/pngloss_b200/exp_profile.so timeout 300 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1 --bm 0,1 --profile > gpurun_out/sweep_y_profile.log 2>&1; cut -c1-250 gpurun_out/sweep_y_profile.log
echo "== bench default"
timeout 900 python bench.py > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_y.json; tail -3 gpurun_out/bench_y.err
