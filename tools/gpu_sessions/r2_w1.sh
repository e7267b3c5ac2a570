#!/bin/bash
# round 2, session W1: racecheck of the latency kernel in the every-lane-arrives build; config 5 on one GPU
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import pngloss_b200
from checkers import Oracle, to_bpp
oracle = Oracle()
ctx = pngloss_b200.Context(0)
rng = np.random.default_rng(5)
def run(w, h, n, s, solo, kind):
    ctx.set_solo(solo)
    imgs = []
    for i in range(n):
        if kind == "synth": a = oracle.synth(w, h, 4 + i)
        elif kind == "noise":
            a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8); a[rng.random((h, w)) < 0.2, 3] = 0
        else: a = (rng.integers(0, 6, (h, w, 4)) * 51).astype(np.uint8)
        imgs.append(to_bpp(a, (i % 4) + 1))
    batch = pngloss_b200.Batch(ctx, [w] * n, [h] * n)
    for i, a in enumerate(imgs): batch.upload(i, a)
    batch.run(s, 2); st, _, _ = batch.finish(); assert (st == 0).all()
    assert batch.launch_info()["solo"]
    out = np.zeros((h, w, 4), np.uint8); rf = np.zeros(h, np.uint8)
    for i in range(n):
        batch.download(i, out, rf); ctx.sync()
        px, want = oracle.optimize(imgs[i], s, 2, True)
        assert np.array_equal(out, px) and np.array_equal(rf, want)
    print("ok", (w, h, n, s, solo, kind), flush=True)
    batch.close()
run(100, 9, 5, 20, 1, "synth")
run(70, 7, 4, 20, 1, "noise")
run(64, 6, 4, 63, 1, "few")
run(100, 9, 5, 20, 2, "synth")
run(70, 7, 4, 126, 2, "noise")
PY
PNGLOSS_B200_LIB=$PWD/pngloss_b200/libsolo_allarrive.so timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python /tmp/san.py > gpurun_out/r2w_racecheck_allarrive.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|^ok\|Error" gpurun_out/r2w_racecheck_allarrive.log | head
timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2w_config5_n1.json 2> gpurun_out/r2w_config5_n1.err; echo "config5 n1 rc=$?"; cut -c1-300 gpurun_out/r2w_config5_n1.json; tail -2 gpurun_out/r2w_config5_n1.err
