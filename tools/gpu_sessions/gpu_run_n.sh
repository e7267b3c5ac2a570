#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_n.log
echo "== bench default"; timeout 1500 python bench.py > gpurun_out/bench_n.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'lanes',d['config']['k2_lanes_per_channel'],'cpu',d['cpu_baseline']['value'])"
