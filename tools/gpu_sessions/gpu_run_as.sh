#!/bin/bash
# session AS: final default bench of the round (and the job timeline for profiles/)
mkdir -p gpurun_out
timeout 600 python tools/e2e_trace.py --images 2368 --height 270 --steps 4 > gpurun_out/e2e_trace_as.log 2>&1; tail -3 gpurun_out/e2e_trace_as.log
echo "== bench default"
timeout 1500 python bench.py > gpurun_out/bench_as.json 2> gpurun_out/bench_as.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_as.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'])
PY
tail -3 gpurun_out/bench_as.err
