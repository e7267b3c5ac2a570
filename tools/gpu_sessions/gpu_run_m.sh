#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_m.log
echo "== chunk trace (reduced height)"; PNGLOSS_B200_TRACE=1 timeout 900 python bench.py --no-cpu --height 270 --steps 2 > gpurun_out/bench_m_small.log 2>&1; grep "pngloss_b200\]" gpurun_out/bench_m_small.log | tail -4; tail -1 gpurun_out/bench_m_small.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])"
echo "== bench full"; PNGLOSS_B200_TRACE=1 timeout 1500 python bench.py --no-cpu > gpurun_out/bench_m.log 2>&1; grep "pngloss_b200\]" gpurun_out/bench_m.log | tail -2; tail -1 gpurun_out/bench_m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])"
