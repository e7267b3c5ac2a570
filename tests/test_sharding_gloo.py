"""CPU tests of the multi-GPU host logic: world_size-2 gloo run of the shard assignment and of the one
collective of the path (the 256-bin symbol histogram all-reduce)."""
import os
import socket
import sys

import numpy as np
import pytest

from pngloss_b200.shard import assign_images, shard_seeds


def test_assign_images_balances_and_is_complete():
    sizes = [(3840, 2160)] * 5 + [(1920, 1080)] * 7 + [(64, 64)] * 3
    for world in (1, 2, 4, 8):
        parts = assign_images(sizes, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[i][0] * sizes[i][1] for i in p) for p in parts]
        assert max(loads) - min(loads) <= 3840 * 2160
    same = assign_images([(1920, 1080)] * 1024, 8)
    assert all(len(p) == 128 for p in same)


def test_shard_seeds_are_disjoint():
    seen = set()
    for r in range(8):
        s = shard_seeds(r, 8, 37)
        assert len(s) == 37 and not (seen & set(s))
        seen |= set(s)
    assert seen == set(range(4, 4 + 8 * 37))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist

    from checkers import Oracle
    from pngloss_b200.shard import allreduce_histogram, assign_images, shard_seeds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    # the checker stands in for the device kernels here: this test is about the plumbing
    local = np.zeros(256, np.uint64)
    for seed in shard_seeds(rank, world, 3):
        _, _, tr = oracle.optimize(oracle.synth(24, 10, seed), 20, 2, True, trace=True)
        local += tr["final_frequency"]
    total = allreduce_histogram(local, dist)
    parts = assign_images([(24, 10)] * 6, world)
    q.put((rank, total, parts[rank]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_histogram_allreduce_world2_gloo():
    import torch.multiprocessing as mp

    from checkers import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    oracle = Oracle()
    want = np.zeros(256, np.uint64)
    for seed in range(4, 10):
        _, _, tr = oracle.optimize(oracle.synth(24, 10, seed), 20, 2, True, trace=True)
        want += tr["final_frequency"]
    for rank, total, part in got:
        assert np.array_equal(total, want)
        assert len(part) == 3
    assert int(want.sum()) == 6 * 24 * 10 * 4
