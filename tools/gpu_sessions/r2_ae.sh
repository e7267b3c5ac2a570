#!/bin/bash
# round 2, session AE: ncu launch list of the default bench command (the 2368 launches of the synthetic-image generator excluded)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pl_k[1234]" -c 400 --csv --log-file gpurun_out/r2ae_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2ae_bench_under_ncu.log 2>&1; echo "ncu rc=$?"; grep -c pl_k gpurun_out/r2ae_launches.csv
