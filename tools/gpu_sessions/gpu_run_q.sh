#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_q.log
echo "== profile"; PNGLOSS_B200_LIB=$PWD/pngloss_b200/exp_profile.so timeout 600 python tools/sweep.py --height 135 --images 1184 --lanes 2,1 --reps 0 --profile > gpurun_out/profile_q.log 2>&1; grep busy gpurun_out/profile_q.log
echo "== sweep"; timeout 600 python tools/sweep.py --height 135 --images 592,1184 --lanes 8,4,2,1 > gpurun_out/sweep_q.log 2>&1; cut -c1-150 gpurun_out/sweep_q.log
echo "== bench default"; timeout 1500 python bench.py > gpurun_out/bench_q.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_q.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'lanes',d['config']['k2_lanes_per_channel'],'cpu',d['cpu_baseline']['value'])"
