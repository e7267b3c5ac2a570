#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_t.log
echo "== sweep"; timeout 600 python tools/sweep.py --height 135 --images 148,296,592,1184 --lanes 8,4,2,1 > gpurun_out/sweep_t.log 2>&1; cut -c1-150 gpurun_out/sweep_t.log
