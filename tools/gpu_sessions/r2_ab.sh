#!/bin/bash
# round 2, session AB (2 GPUs): config 4 (1024 x 1080p, strong scaling): 512 images per GPU run on the latency kernel's four-warp layout
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 2 --config 4 --steps 3 --warmup 1 > gpurun_out/r2ab_config4_n2.json 2> gpurun_out/r2ab_config4_n2.err; echo "config4 n2 rc=$?"; tail -1 gpurun_out/r2ab_config4_n2.json | cut -c1-300; tail -2 gpurun_out/r2ab_config4_n2.err
