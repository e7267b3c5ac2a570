#!/bin/bash
# session AI (2 GPUs): the bench under torchrun, reduced height (270 rows) to save GPU minutes
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --height 270 > gpurun_out/bench_ai_2gpu.json 2> gpurun_out/bench_ai_2gpu.err; echo "rc=$?"; cat gpurun_out/bench_ai_2gpu.json | cut -c1-1800; tail -5 gpurun_out/bench_ai_2gpu.err
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --height 270 --no-cpu > gpurun_out/bench_ai_1gpu.json 2> gpurun_out/bench_ai_1gpu.err; echo "rc=$?"; cat gpurun_out/bench_ai_1gpu.json | cut -c1-900
