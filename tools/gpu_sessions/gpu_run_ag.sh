#!/bin/bash
# session AG: job API timeline with merged copies; parity of the host-buffer paths; bench
mkdir -p gpurun_out
echo "== e2e trace"
timeout 600 python tools/e2e_trace.py 2>&1 | tee gpurun_out/e2e_trace_ag.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_ag.log
echo "== bench default"
timeout 1200 python bench.py > gpurun_out/bench_ag.json 2> gpurun_out/bench_ag.err; echo "bench rc=$?"; cat gpurun_out/bench_ag.json | cut -c1-1500; tail -5 gpurun_out/bench_ag.err
