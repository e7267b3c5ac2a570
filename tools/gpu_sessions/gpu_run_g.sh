#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_g.log
echo "== bench default"; timeout 1500 python bench.py > gpurun_out/bench_g.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_g.log | cut -c1-1200
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_g.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref_g.log | cut -c1-600
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pl_k[123] -c 30 --csv --log-file gpurun_out/launches_g.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_g.log 2>&1; echo "rc=$?"
echo "== ncu dram bytes"; timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pl_k[12] -s 2 -c 2 --csv --log-file gpurun_out/dram_g.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_dram_g.log 2>&1; echo "rc=$?"; cat gpurun_out/dram_g.csv | tail -8
