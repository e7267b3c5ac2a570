#!/bin/bash
# round 2, session S: whole GPU suite with the latency kernel as the default for small batches, and the single-image
# bench lines (configs 1, 2, 3 at four strengths) again
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest.log
for c in 1 2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2s_config$c.json 2> gpurun_out/r2s_config$c.err; echo "config $c rc=$?"; cut -c1-200 gpurun_out/r2s_config$c.json
done
for s in 20 0 40 85; do
  timeout 400 python bench.py --config 3 --strength $s --steps 2 --warmup 1 > gpurun_out/r2s_config3_s$s.json 2> gpurun_out/r2s_config3_s$s.err; echo "config 3 s$s rc=$?"; cut -c1-200 gpurun_out/r2s_config3_s$s.json
done
