#!/bin/bash
# round 2, session J: compute-sanitizer (memcheck + racecheck) on the final K1 / K2 (generic and lean) / K3 / K4,
# ncu of K1 and K2 on lena (BASELINE configs[1]), new K1-vs-reference test
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import pngloss_b200
from checkers import Oracle, to_bpp
oracle = Oracle()
ctx = pngloss_b200.Context(0)
def run(w, h, n, s, lanes, bm, lean, modes=False):
    ctx.set_lanes(lanes); ctx.set_bucket_maxima(bm); ctx.set_lean(lean)
    imgs = [to_bpp(oracle.synth(w, h, 4 + i), (i % 4) + 1 if modes else 4) for i in range(n)]
    batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
    for i, a in enumerate(imgs): batch.upload(i, a)
    batch.run(s, 2); st, _, _ = batch.finish(); assert (st == 0).all()
    batch.scanlines(); batch.scanline_info(0)
    out = np.zeros((h, w, 4), np.uint8); rf = np.zeros(h, np.uint8)
    batch.download(0, out, rf); ctx.sync()
    px, want = oracle.optimize(imgs[0], s, 2, True)
    assert np.array_equal(out, px) and np.array_equal(rf, want)
    print("ok", (w, h, n, s, lanes, bm, lean), batch.launch_info(), flush=True)
    batch.close()
run(64, 12, 8, 20, 1, 1, 1)           # lean, all lanes active
run(100, 9, 11, 20, 1, 1, 1, True)    # lean, mixed modes, ragged CTA
run(64, 12, 8, 20, 1, 1, 0)           # generic, bucket maxima
run(61, 9, 3, 20, 8, 0, 0, True)      # generic, wide lanes, scan
run(64, 8, 5, 255, 2, 0, 0)
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python /tmp/san.py > gpurun_out/r2j_memcheck.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|^ok" gpurun_out/r2j_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python /tmp/san.py > gpurun_out/r2j_racecheck.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|^ok" gpurun_out/r2j_racecheck.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k1_original" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pl_k1|pl_k2" -c 2 -f -o gpurun_out/r2j_lena python bench.py --config 2 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2j_ncu_lena.log 2>&1; echo "ncu rc=$?"
