#!/bin/bash
# round 2, session AA: the four-warp layout of the latency kernel (up to four CTAs per SM) against the eight-warp one
# and against the generic kernel's lane mappings
mkdir -p gpurun_out
timeout 300 python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import pngloss_b200
from checkers import Oracle, to_bpp
oracle = Oracle(); ctx = pngloss_b200.Context(0); ctx.set_solo(3)
rng = np.random.default_rng(3)
for (w, h, n, s) in [(100, 9, 7, 20), (70, 12, 5, 63), (33, 20, 6, 5)]:
    imgs = [to_bpp(oracle.synth(w, h, 4 + i) if i % 2 else rng.integers(0, 256, (h, w, 4), dtype=np.uint8), (i % 4) + 1) for i in range(n)]
    batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
    for i, a in enumerate(imgs): batch.upload(i, a)
    batch.run(s, 2); st, _, _ = batch.finish(); assert (st == 0).all() and batch.launch_info()["solo"]
    out = np.zeros((h, w, 4), np.uint8); rf = np.zeros(h, np.uint8)
    for i in range(n):
        batch.download(i, out, rf); ctx.sync()
        px, want = oracle.optimize(imgs[i], s, 2, True)
        assert np.array_equal(out, px) and np.array_equal(rf, want), (w, h, s, i)
    batch.close()
print("four-warp layout: parity ok")
PY
{
timeout 600 python tools/sweep.py --width 3840 --height 135 --images 1,148,296,444,592 --lanes 0 --solo 3 --reps 1
timeout 600 python tools/sweep.py --width 3840 --height 135 --images 1,148,296 --lanes 0 --solo 1 --reps 1
timeout 600 python tools/sweep.py --width 3840 --height 135 --images 444,592 --lanes 0 --solo 0 --reps 1
} 2>&1 | cut -c1-260 > gpurun_out/r2aa_sweep.txt
cat gpurun_out/r2aa_sweep.txt
