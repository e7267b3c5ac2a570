#!/bin/bash
# session AK: K4 (filtered scanlines): parity, device time
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ak.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_ak.log
echo "== k4 bench (K2 runs once on 270 rows first)"
timeout 600 python tools/k4_bench.py --images 1184 --height 270 2>&1 | tail -2
timeout 600 python tools/k4_bench.py --images 296 --height 2160 2>&1 | tail -2
