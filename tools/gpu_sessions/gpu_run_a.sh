#!/bin/bash
# first GPU session: parity, sanitizer, sweep, bench, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sanitizer"; 
timeout 300 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout 400 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
echo "== sweep small"; timeout 300 python tools/sweep.py --height 270 --images 148,296,592,1184 --lanes 8,4,2,1 > gpurun_out/sweep_small.log 2>&1; cat gpurun_out/sweep_small.log
echo "== sweep 4k"; timeout 600 python tools/sweep.py --images 148,296 --lanes 8,4,2 --reps 0 > gpurun_out/sweep_4k.log 2>&1; cat gpurun_out/sweep_4k.log
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 3 --images 148 > gpurun_out/bench_a.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench_a.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --images 148 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pl_k -c 4 -o gpurun_out/prof_a python tools/sweep.py --height 64 --images 296 --lanes 8 --reps 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
