#!/bin/bash
# round 2, session H: which end-to-end set-up for the default bench line?  pipeline of 296-image jobs (lean kernel,
# 3552 images per step) against two 2368-image jobs in flight (generic kernel), both over 8 steps; synth kernel speed
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r2h_3552_pipeline.json 2> gpurun_out/r2h_3552_pipeline.err; echo "pipeline rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_3552_pipeline.json").read().strip().splitlines()[-1]); print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],d["kernel_ms"])
PY
timeout 900 python bench.py --images 2368 --lean 0 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2h_2368_twojobs.json 2> gpurun_out/r2h_2368_twojobs.err; echo "two jobs rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_2368_twojobs.json").read().strip().splitlines()[-1]); print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],d["kernel_ms"])
PY
