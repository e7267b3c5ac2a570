#!/bin/bash
# session AB: does K2's speed depend on the distance between the images' buffers (DRAM / L2 camping)?
mkdir -p gpurun_out
for pad in 0 256 768 2048 4352 15360 33024 66304 1048832; do
  echo "== pad $pad"
  PNGLOSS_B200_IMAGE_PAD=$pad timeout 300 python tools/sweep.py --height 135 --images 1184 --lanes 1 --bm 1 2>&1 | cut -c1-150
done
