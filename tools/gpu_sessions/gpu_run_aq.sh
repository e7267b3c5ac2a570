#!/bin/bash
# session AQ: default bench with the leaner host buffers of the end-to-end part; GPU suite incl. the new width test
mkdir -p gpurun_out
free -g | head -2
echo "== bench default"
timeout 1500 python bench.py --no-cpu > gpurun_out/bench_aq.json 2> gpurun_out/bench_aq.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_aq.json'))
print(d['value'], d['e2e'])
PY
tail -3 gpurun_out/bench_aq.err
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_aq.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_aq.log
