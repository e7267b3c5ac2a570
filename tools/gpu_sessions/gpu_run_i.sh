#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_i.log
echo "== bench default"; timeout 1500 python bench.py --no-cpu > gpurun_out/bench_i.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_i.log | cut -c1-1300
