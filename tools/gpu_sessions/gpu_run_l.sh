#!/bin/bash
mkdir -p gpurun_out
echo "== chunk trace (reduced height)"; PNGLOSS_B200_TRACE=1 timeout 900 python bench.py --no-cpu --height 270 --steps 2 > gpurun_out/bench_l.log 2>&1; grep "pngloss_b200\]" gpurun_out/bench_l.log | tail -8; tail -1 gpurun_out/bench_l.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])"
