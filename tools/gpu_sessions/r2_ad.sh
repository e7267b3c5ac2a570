#!/bin/bash
# round 2, session AD: final build - ncu launch list of the default bench command, then the default bench as the driver runs it
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ad_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2ad_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ad_default.json 2> gpurun_out/r2ad_default.err; echo "default rc=$?"; cut -c1-300 gpurun_out/r2ad_default.json; tail -2 gpurun_out/r2ad_default.err
