#!/bin/bash
# session AR: job timeline at 2368 images (270 and 540 rows)
mkdir -p gpurun_out
timeout 600 python tools/e2e_trace.py --images 2368 --height 270 --steps 4 2>&1 | tail -16
timeout 600 python tools/e2e_trace.py --images 2368 --height 540 --steps 3 2>&1 | tail -12
