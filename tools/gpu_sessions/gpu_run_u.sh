#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_u.log
