"""GPU tests of the multi-GPU part of the C ABI (pl_comm.cuh): the path's one collective - the NCCL sum of the
batch symbol histograms - issued by the library itself, without torch.  The 1-rank cases run on any GPU box;
the 2-rank cases need two GPUs (gpurun --gpus 2) and are skipped otherwise."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import pngloss_b200
from checkers import Oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def expected_total(oracle, world):
    total = np.zeros(256, np.uint64)
    for rank in range(world):
        for i in range(3 + rank):
            _, _, tr = oracle.optimize(oracle.synth(40, 12, 1000 * rank + i), 20, 2, True, trace=True)
            total += tr["final_frequency"].astype(np.uint64)
    return total


def test_single_rank_communicator():
    """nranks = 1: the all-reduce is the identity, but the whole path (dlopen of libnccl.so.2, communicator on
    the context's device, collective on the context's stream) runs."""
    oracle = Oracle()
    ctx = pngloss_b200.Context(0)
    ctx.comm_init_rank(1, 0, pngloss_b200.comm_unique_id())
    assert ctx.comm_size() == 1
    imgs = [oracle.synth(40, 12, i) for i in range(3)]
    batch = pngloss_b200.Batch(ctx, [40] * 3, [12] * 3)
    for i, a in enumerate(imgs):
        batch.upload(i, a)
    batch.run(20, 2)
    batch.allreduce_histogram()
    st, _, _ = batch.finish()
    assert (st == 0).all()
    assert np.array_equal(batch.histogram(), expected_total(oracle, 1))
    assert ctx.comm_allreduce([7, 9], "max").tolist() == [7, 9]
    ctx.barrier()
    batch.close()
    ctx.close()


@pytest.mark.skipif(pngloss_b200.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_two_processes(tmp_path):
    """One process per GPU (the torchrun layout), no torch imported: the reduced histogram on every rank is the
    sum of the oracle's per-image histograms over all ranks."""
    oracle = Oracle()
    id_file = str(tmp_path / "nccl.id")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "comm_worker.py"), str(r), "2", id_file,
                               str(tmp_path / f"out{r}.npz")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "torch" not in "".join(outs)
    want = expected_total(oracle, 2)
    for r in range(2):
        z = np.load(str(tmp_path / f"out{r}.npz"))
        assert np.array_equal(z["hist"], want), r
        assert np.array_equal(z["cum"], want), r
        assert int(z["tmax"]) == 6 and z["tsum"].tolist() == [11, 2]


@pytest.mark.skipif(pngloss_b200.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_one_process_threads():
    """One process, one context and host thread per GPU (what the command line's --gpus N does)."""
    oracle = Oracle()
    ctxs = [pngloss_b200.Context(d) for d in range(2)]
    pngloss_b200.comm_init_all(ctxs)
    got = [None, None]

    def work(rank):
        ctx = ctxs[rank]
        n = 3 + rank
        imgs = [oracle.synth(40, 12, 1000 * rank + i) for i in range(n)]
        batch = pngloss_b200.Batch(ctx, [40] * n, [12] * n)
        for i, a in enumerate(imgs):
            batch.upload(i, a)
        batch.run(20, 2)
        batch.allreduce_histogram()
        batch.finish()
        got[rank] = batch.histogram()
        batch.close()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    want = expected_total(oracle, 2)
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)
    for c in ctxs:
        c.close()
