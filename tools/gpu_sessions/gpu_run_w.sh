#!/bin/bash
# session W: ncu --set full of K2 <1 lane per channel, bucket maxima>, 1184 images (1 CTA/SM) and 2368 images (2 CTAs/SM)
mkdir -p gpurun_out
for n in 1184 2368; do
  echo "== ncu K2 lanes 1 bm 1 images $n"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pl_k2 -c 1 -f -o gpurun_out/prof_w_k2_l1bm_$n python tools/sweep.py --height 24 --images $n --lanes 1 --bm 1 --reps 0 > gpurun_out/ncu_w_$n.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_w_$n.log
done
ls -la gpurun_out
