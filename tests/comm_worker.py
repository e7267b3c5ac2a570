"""Helper of tests/test_gpu_comm.py: one rank of a multi-process run that uses the library's own NCCL
communicator (no torch).  argv: rank world id_file out_file"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pngloss_b200  # noqa: E402
from checkers import Oracle  # noqa: E402


def main():
    rank, world, id_file, out_file = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    if rank == 0:
        uid = pngloss_b200.comm_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        uid = open(id_file, "rb").read()
    ctx = pngloss_b200.Context(rank)
    ctx.comm_init_rank(world, rank, uid)
    oracle = Oracle()
    n = 3 + rank                                      # ranks hold different numbers of different images
    imgs = [oracle.synth(40, 12, 1000 * rank + i) for i in range(n)]
    batch = pngloss_b200.Batch(ctx, [40] * n, [12] * n)
    for i, a in enumerate(imgs):
        batch.upload(i, a)
    batch.run(20, 2)
    batch.allreduce_histogram()                       # the path's one collective, issued by the library
    st, _, _ = batch.finish()
    assert (st == 0).all()
    hist = batch.histogram()
    tmax = ctx.comm_allreduce([rank + 5], "max")[0]
    tsum = ctx.comm_allreduce([rank + 5, 1], "sum")
    # host-buffer path + cumulative symbol histogram across ranks
    work = [a.copy() for a in imgs]
    res = ctx.optimize_batch(work, [np.zeros(12, np.uint8) for _ in work], 20, 2)
    assert all(r["status"] == 0 for r in res)
    cum = ctx.symbol_histogram(across_ranks=True)
    np.savez(out_file, hist=hist, tmax=tmax, tsum=tsum, cum=cum)
    ctx.barrier()
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
