#!/bin/bash
# session AV: full GPU suite with the pool / grouped-with-scanlines test
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_av.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_av.log
