#!/bin/bash
# round 2, session U: A/B of the latency kernel's fast-path coverage (one-value bands, in-lane scan) on real images
mkdir -p gpurun_out
for v in 00 10 11; do
  echo "== single/fastscan = $v"
  PNGLOSS_B200_LIB=$PWD/pngloss_b200/libsolo$v.so timeout 300 python tools/solo_ab.py
done > gpurun_out/r2u_variants.txt 2>&1
cat gpurun_out/r2u_variants.txt
