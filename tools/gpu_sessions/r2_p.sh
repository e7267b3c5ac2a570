#!/bin/bash
# round 2, session P: ncu source-level profile of the latency kernel's chain loop (one 3840 x 64 image)
mkdir -p gpurun_out
for so in 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pl_k2_solo -c 1 -f -o gpurun_out/r2p_solo$so python tools/sweep.py --width 3840 --height 64 --images 1 --lanes 8 --solo $so --reps 0 > gpurun_out/r2p_ncu$so.log 2>&1; echo "ncu rc=$?"
done
