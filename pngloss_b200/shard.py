"""Host-side sharding of independent images over GPUs.

Images are the shard unit: the quantise path has no cross-image dependency, so there is no data-path
collective.  The only cross-rank quantity is the batch symbol histogram (256 x u64), summed with one
all-reduce.  Pure Python so that the logic is testable on CPU with the gloo backend.
"""
from typing import List, Sequence, Tuple

import numpy as np


def assign_images(sizes: Sequence[Tuple[int, int]], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment by pixel count: image indices per rank.
    Equal sizes degenerate to round-robin; deterministic for a given input order."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i][0] * sizes[i][1], i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i][0] * sizes[i][1]
    for lst in out:
        lst.sort()
    return out


def shard_seeds(rank: int, world: int, images_per_rank: int, first_seed: int = 4) -> List[int]:
    """bench.py's weak-scaling workload: every rank owns `images_per_rank` distinct synthetic images."""
    assert 0 <= rank < world
    return [first_seed + rank * images_per_rank + i for i in range(images_per_rank)]


def allreduce_histogram(local_hist: np.ndarray, dist=None, device=None) -> np.ndarray:
    """Sum a (256,) uint64 histogram over all ranks.  Counts of a 1024 x 1080p batch exceed 2^32, hence
    64-bit.  `dist` is torch.distributed (already initialised) or None for a single rank."""
    assert local_hist.shape == (256,)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_hist.astype(np.uint64)
    import torch
    t = torch.from_numpy(local_hist.astype(np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy().astype(np.uint64)
