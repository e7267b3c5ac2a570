#!/bin/bash
# round 2, session AF: single-image bench lines of the final build
mkdir -p gpurun_out
for c in 1 2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2af_config$c.json 2> gpurun_out/r2af_config$c.err; echo "config $c rc=$?"; cut -c1-160 gpurun_out/r2af_config$c.json
done
for s in 0 20 40 85; do
  timeout 400 python bench.py --config 3 --strength $s --steps 2 --warmup 1 > gpurun_out/r2af_config3_s$s.json 2> gpurun_out/r2af_config3_s$s.err; echo "config 3 s$s rc=$?"; cut -c1-160 gpurun_out/r2af_config3_s$s.json
done
