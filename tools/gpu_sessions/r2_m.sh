#!/bin/bash
# round 2, session M (8 GPUs): the torch-free multi-rank bench at N=8 - default config (weak), config 4 and 5 (strong)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2m_box.txt; free -g >> gpurun_out/r2m_box.txt; nproc >> gpurun_out/r2m_box.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
PNGLOSS_BENCH_TRACE=1 timeout 900 $TR bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2m_config3s_n8.json 2> gpurun_out/r2m_config3s_n8.err; echo "3s n8 rc=$?"; tail -1 gpurun_out/r2m_config3s_n8.json | cut -c1-300
timeout 600 $TR bench.py --gpus 8 --config 4 --steps 2 --warmup 1 > gpurun_out/r2m_config4_n8.json 2> gpurun_out/r2m_config4_n8.err; echo "config4 n8 rc=$?"; tail -1 gpurun_out/r2m_config4_n8.json | cut -c1-300
timeout 600 $TR bench.py --gpus 8 --config 5 --steps 1 --warmup 0 --no-e2e > gpurun_out/r2m_config5_n8.json 2> gpurun_out/r2m_config5_n8.err; echo "config5 n8 rc=$?"; tail -1 gpurun_out/r2m_config5_n8.json | cut -c1-300
grep "e2e-trace" gpurun_out/r2m_config3s_n8.err | sort -k3,3n -k5,5n | head -40
