#!/bin/bash
# session AM: ncu --set full of the K4 filter kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k4_scanlines -c 1 -f -o gpurun_out/prof_am_k4 python tools/k4_bench.py --images 296 --height 540 > gpurun_out/ncu_am_k4.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_am_k4.log
