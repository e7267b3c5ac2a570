#!/bin/bash
# round 2, session AI: the committed final build once more - whole GPU suite, smoke(), single-image lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ai_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for c in 1 2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2ai_config$c.json 2> gpurun_out/r2ai_config$c.err; echo "config $c rc=$?"; cut -c1-160 gpurun_out/r2ai_config$c.json
done
for s in 0 20 40 85; do
  timeout 400 python bench.py --config 3 --strength $s --steps 2 --warmup 1 --no-cpu > gpurun_out/r2ai_config3_s$s.json 2> gpurun_out/r2ai_config3_s$s.err; echo "config 3 s$s rc=$?"; cut -c1-160 gpurun_out/r2ai_config3_s$s.json
done
