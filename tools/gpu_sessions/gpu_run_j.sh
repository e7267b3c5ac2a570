#!/bin/bash
mkdir -p gpurun_out
uptime > gpurun_out/host_j.txt; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,utilization.gpu,memory.used --format=csv >> gpurun_out/host_j.txt
echo "== bench small trace"; PNGLOSS_BENCH_TRACE=1 timeout 600 python bench.py --no-cpu --images 592 --height 540 > gpurun_out/bench_j_small.log 2>&1; grep trace gpurun_out/bench_j_small.log; tail -1 gpurun_out/bench_j_small.log | cut -c1-400
echo "== bench full trace"; PNGLOSS_BENCH_TRACE=1 timeout 1500 python bench.py --no-cpu > gpurun_out/bench_j.log 2>&1; grep trace gpurun_out/bench_j.log; tail -1 gpurun_out/bench_j.log | cut -c1-1300
uptime >> gpurun_out/host_j.txt; cat gpurun_out/host_j.txt
