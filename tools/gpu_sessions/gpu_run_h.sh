#!/bin/bash
# 2-GPU validation of the torchrun path
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_h.txt; free -g | head -2 >> gpurun_out/gpus_h.txt; nproc >> gpurun_out/gpus_h.txt
echo "== torchrun 2 ranks (reduced height)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --images 592 --height 540 > gpurun_out/bench_h_2gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_h_2gpu.log | cut -c1-1500
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-rows 64 > gpurun_out/bench_h_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_h_ref.log | cut -c1-400
echo "== single gpu same workload for comparison"; timeout 600 python bench.py --steps 2 --warmup 3 --images 592 --height 540 --no-cpu > gpurun_out/bench_h_1gpu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_h_1gpu.log | cut -c1-700
