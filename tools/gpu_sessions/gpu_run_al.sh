#!/bin/bash
# session AL: vectorised K4: parity + device time
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "k4" 2>&1 | tail -3
timeout 600 python tools/k4_bench.py --images 1184 --height 270 2>&1 | tail -2
timeout 600 python tools/k4_bench.py --images 296 --height 2160 2>&1 | tail -2
