#!/bin/bash
# round 2, session V: final build - whole GPU suite, compute-sanitizer on the latency kernel, ncu of K1 / K2S on lena,
# the single-image bench lines, the default bench as the driver runs it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
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
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python /tmp/san.py > gpurun_out/r2v_memcheck.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|^ok" gpurun_out/r2v_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 10 python /tmp/san.py > gpurun_out/r2v_racecheck.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|^ok" gpurun_out/r2v_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pl_k1|pl_k2" -c 2 -f -o gpurun_out/r2v_lena python bench.py --config 2 --steps 1 --warmup 0 --no-e2e --no-cpu --no-k4 > gpurun_out/r2v_ncu_lena.log 2>&1; echo "ncu rc=$?"
for c in 1 2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2v_config$c.json 2> gpurun_out/r2v_config$c.err; echo "config $c rc=$?"; cut -c1-160 gpurun_out/r2v_config$c.json
done
for s in 20 40 85; do
  timeout 400 python bench.py --config 3 --strength $s --steps 2 --warmup 1 > gpurun_out/r2v_config3_s$s.json 2> gpurun_out/r2v_config3_s$s.err; echo "config 3 s$s rc=$?"; cut -c1-160 gpurun_out/r2v_config3_s$s.json
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_default.json 2> gpurun_out/r2v_default.err; echo "default rc=$?"; cut -c1-400 gpurun_out/r2v_default.json; tail -3 gpurun_out/r2v_default.err
