#!/bin/bash
# session AF: job API timeline after moving the row filters to pinned staging; parity; bench
mkdir -p gpurun_out
echo "== e2e trace"
timeout 600 python tools/e2e_trace.py 2>&1 | tee gpurun_out/e2e_trace_af.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_af.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_af.log
echo "== bench default"
timeout 1200 python bench.py > gpurun_out/bench_af.json 2> gpurun_out/bench_af.err; echo "bench rc=$?"; cat gpurun_out/bench_af.json | cut -c1-1500; tail -5 gpurun_out/bench_af.err
