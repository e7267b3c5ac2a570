#!/bin/bash
# round 2, session G: whole GPU suite (reference main, comm, CLI fixes, suite + 8192 goldens) and the first
# per-config bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
for c in 1 2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2g_config$c.json 2> gpurun_out/r2g_config$c.err; echo "config $c rc=$?"; cut -c1-300 gpurun_out/r2g_config$c.json
done
for s in 20 0 40 85; do
  timeout 400 python bench.py --config 3 --strength $s --steps 2 --warmup 1 > gpurun_out/r2g_config3_s$s.json 2> gpurun_out/r2g_config3_s$s.err; echo "config 3 s$s rc=$?"; cut -c1-300 gpurun_out/r2g_config3_s$s.json
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_config3s.json 2> gpurun_out/r2g_config3s.err; echo "config 3s rc=$?"; cat gpurun_out/r2g_config3s.json; tail -5 gpurun_out/r2g_config3s.err
