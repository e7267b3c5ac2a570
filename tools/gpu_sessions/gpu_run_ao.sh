#!/bin/bash
# session AO: scanlines through the host-buffer call and the command line; full GPU suite
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ao.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_ao.log
