#!/bin/bash
# round 2, session N: K4 (filtered scanlines) with 1 / 2 / 4 four-pixel groups per thread; ncu of the default
mkdir -p gpurun_out
for v in k4g1 "" k4g4; do
  lib=pngloss_b200/libpngloss_b200.so; [ -n "$v" ] && lib=pngloss_b200/lib$v.so
  echo "== ${v:-default (2 groups)}"
  PNGLOSS_B200_LIB=$PWD/$lib timeout 300 python tools/k4_bench.py --images 1184 --height 540
done > gpurun_out/r2n_k4.txt 2>&1
cat gpurun_out/r2n_k4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k4_scanlines -c 1 -f -o gpurun_out/r2n_k4 python tools/k4_bench.py --images 1184 --height 540 > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -m gpu -k "k4 or scanline or cli or dropin or main" 2>&1 | tail -2
