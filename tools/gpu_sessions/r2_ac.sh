#!/bin/bash
# round 2, session AC: final build - whole GPU suite, smoke(), a short default bench (contract check)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ac_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2ac_default_short.json 2> gpurun_out/r2ac_default_short.err; echo "default rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2ac_default_short.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "k4", d["roofline_k4"]["frac"], "k1", d["roofline_k1"]["frac"], d["config"]["k2_kernel"], d["clocks"])
PY
tail -2 gpurun_out/r2ac_default_short.err
