#!/bin/bash
# round 2, session I (2 GPUs): the library's NCCL path (no torch), CLI --gpus 2, bench under torchrun at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2i_gpus.txt; free -g >> gpurun_out/r2i_gpus.txt; nproc >> gpurun_out/r2i_gpus.txt
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_cli.py -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --config 4 --steps 2 --warmup 1 > gpurun_out/r2i_config4_n2.json 2> gpurun_out/r2i_config4_n2.err; echo "config4 n2 rc=$?"; tail -1 gpurun_out/r2i_config4_n2.json | cut -c1-400
timeout 600 python bench.py --config 4 --steps 2 --warmup 1 > gpurun_out/r2i_config4_n1.json 2> gpurun_out/r2i_config4_n1.err; echo "config4 n1 rc=$?"; tail -1 gpurun_out/r2i_config4_n1.json | cut -c1-400
timeout 900 $TR bench.py --gpus 2 --config 5 --steps 1 --warmup 1 > gpurun_out/r2i_config5_n2.json 2> gpurun_out/r2i_config5_n2.err; echo "config5 n2 rc=$?"; tail -1 gpurun_out/r2i_config5_n2.json | cut -c1-400
timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2i_config3s_n2.json 2> gpurun_out/r2i_config3s_n2.err; echo "3s n2 rc=$?"; tail -1 gpurun_out/r2i_config3s_n2.json | cut -c1-600
grep -l torch /proc/*/maps 2>/dev/null | head -1
tail -3 gpurun_out/r2i_*.err
