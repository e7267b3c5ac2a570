#!/bin/bash
# session AP: smoke(), the default bench and its reference arm, ncu launch list + DRAM bytes of the bench command
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench default"
timeout 1500 python bench.py > gpurun_out/bench_ap.json 2> gpurun_out/bench_ap.err; echo "bench rc=$?"; cat gpurun_out/bench_ap.json | cut -c1-1200; tail -3 gpurun_out/bench_ap.err
echo "== ncu launch list + dram bytes of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pl_k[123] --csv --log-file gpurun_out/launches_ap.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_ap_bench.log 2>&1; echo "rc=$?"; grep -c pl_k gpurun_out/launches_ap.csv
