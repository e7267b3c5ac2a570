#!/bin/bash
# session V: bucket-maxima variant of K2 - parity, lane x variant sweep, default bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_v.log
echo "== sweep 1184/2368 lanes 1,2 bm 0,1"
timeout 600 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1,2 --bm 0,1 > gpurun_out/sweep_v_a.log 2>&1; cut -c1-170 gpurun_out/sweep_v_a.log
echo "== sweep 148..592 lanes all bm 0,1"
timeout 600 python tools/sweep.py --height 135 --images 148,296,592 --lanes 8,4,2,1 --bm 0,1 > gpurun_out/sweep_v_b.log 2>&1; cut -c1-170 gpurun_out/sweep_v_b.log
echo "== bench default"
timeout 900 python bench.py > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"; cat gpurun_out/bench_v.json; tail -3 gpurun_out/bench_v.err
