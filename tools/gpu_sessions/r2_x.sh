#!/bin/bash
# round 2, session X: K4 with warp-contiguous groups and aligned copy-out reads: time, parity tests, ncu
mkdir -p gpurun_out
timeout 300 python tools/k4_bench.py --images 1184 --height 540 > gpurun_out/r2x_k4.txt 2>&1; cat gpurun_out/r2x_k4.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -m gpu -k "k4 or scanline or cli or dropin or main" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k4_scanlines -c 1 -f -o gpurun_out/r2x_k4 python tools/k4_bench.py --images 1184 --height 540 > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc=$?"
