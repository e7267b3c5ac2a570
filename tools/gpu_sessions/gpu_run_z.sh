#!/bin/bash
# session Z: spare-warp scheduler balancing, in-place batches, job API; parity, A/B sweep, bench
mkdir -p gpurun_out
free -g | head -2; nproc
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_z.log
echo "== sweep with spare warp"
timeout 600 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1 --bm 1 > gpurun_out/sweep_z_spare.log 2>&1; cut -c1-170 gpurun_out/sweep_z_spare.log
echo "== sweep without spare warp"
PNGLOSS_B200_NO_SPARE_WARP=1 timeout 600 python tools/sweep.py --height 135 --images 1184,2368 --lanes 1 --bm 1 > gpurun_out/sweep_z_nospare.log 2>&1; cut -c1-170 gpurun_out/sweep_z_nospare.log
echo "== per-filter busy with spare warp"
PNGLOSS_B200_LIB=$PWD/pngloss_b200/exp_profile.so timeout 300 python tools/sweep.py --height 135 --images 2368 --lanes 1 --bm 1 --profile > gpurun_out/sweep_z_profile.log 2>&1; cut -c1-250 gpurun_out/sweep_z_profile.log
echo "== bench default"
timeout 1200 python bench.py > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench rc=$?"; cat gpurun_out/bench_z.json; tail -5 gpurun_out/bench_z.err
