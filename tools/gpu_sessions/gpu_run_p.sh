#!/bin/bash
mkdir -p gpurun_out
echo "== bench 2368 images (2 CTAs/SM at 8 images per CTA), device-resident only"
timeout 1200 python bench.py --images 2368 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_p_2368.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_p_2368.log | cut -c1-900
echo "== sweep h=135 for 2368"; timeout 600 python tools/sweep.py --height 135 --images 2368 --lanes 2,1 > gpurun_out/sweep_p.log 2>&1; cut -c1-150 gpurun_out/sweep_p.log
