#!/bin/bash
# session AD: spare warp removed, batched row commit; parity, sweep, bench (pipelined e2e)
mkdir -p gpurun_out
echo "== sweep"
timeout 600 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1,2 --bm 1 > gpurun_out/sweep_ad.log 2>&1; cut -c1-170 gpurun_out/sweep_ad.log
timeout 600 python tools/sweep.py --height 135 --images 148,592 --lanes 8,4,2 --bm 0 >> gpurun_out/sweep_ad.log 2>&1; tail -6 gpurun_out/sweep_ad.log | cut -c1-170
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ad.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_ad.log
echo "== bench default"
timeout 1200 python bench.py > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench rc=$?"; cat gpurun_out/bench_ad.json | cut -c1-1900; tail -5 gpurun_out/bench_ad.err
