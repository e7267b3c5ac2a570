#!/bin/bash
# session AH: default bench at 2368 images; ncu launch list + DRAM bytes of the bench command; ncu --set full of K2
mkdir -p gpurun_out
echo "== bench default"
timeout 1500 python bench.py > gpurun_out/bench_ah.json 2> gpurun_out/bench_ah.err; echo "bench rc=$?"; cat gpurun_out/bench_ah.json | cut -c1-1700; tail -5 gpurun_out/bench_ah.err
echo "== reference arm"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ah_ref.json 2>> gpurun_out/bench_ah.err; cat gpurun_out/bench_ah_ref.json | cut -c1-600
echo "== ncu launch list + dram bytes of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pl_k[123] --csv --log-file gpurun_out/launches_ah.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_ah_bench.log 2>&1; echo "rc=$?"; grep -v "^==" gpurun_out/launches_ah.csv | cut -d, -f5,13- | head -30
echo "== ncu --set full K2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_ah_k2_l1bm_2368 python tools/sweep.py --height 24 --images 2368 --lanes 1 --bm 1 --reps 0 > gpurun_out/ncu_ah_k2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_ah_k2.log
