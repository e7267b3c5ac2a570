#!/bin/bash
# session AU (2 GPUs): the final bench.py under torchrun, reduced height; and its reference arm under torchrun
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --height 270 --no-cpu > gpurun_out/bench_au_2gpu.json 2> gpurun_out/bench_au_2gpu.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_au_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['e2e'], d['config']['collective'])
PY
tail -3 gpurun_out/bench_au_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-rows 64 2>/dev/null | cut -c1-300
