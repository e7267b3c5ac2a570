#!/bin/bash
# round 2, session W2 (8 GPUs): config 4 (1024 x 1080p, strong scaling) with the latency kernel: 128 images per GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --config 4 --steps 3 --warmup 1 > gpurun_out/r2w_config4_n8.json 2> gpurun_out/r2w_config4_n8.err; echo "config4 n8 rc=$?"; tail -1 gpurun_out/r2w_config4_n8.json | cut -c1-400; tail -3 gpurun_out/r2w_config4_n8.err
