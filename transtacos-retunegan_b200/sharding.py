"""Multi-GPU plumbing: utterances shard across ranks with NO data-path collective.

One process per GPU (torch.distributed, NCCL over NVLink on the box, gloo in CPU tests).  Feature extraction
and Griffin-Lim have no cross-utterance term, so a rank just takes its utterances; the only exchanges are an
optional min/max reduce of corpus statistics (transtacos/datasets/databaker.py:82-87,118-123) and the scalar
loss mean of ``multi_stft_loss(ddp_reduce=True)``.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-balanced assignment: sort by length descending, give each utterance to the least-loaded rank.
    Deterministic (ties -> lowest rank / lowest index).  Returns per-rank lists of utterance indices."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    lengths = np.asarray(lengths, np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, np.int64)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        shards[r].append(int(i))
        load[r] += int(lengths[i])
    return shards


def rank_world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def reduce_stats(local_min: float, local_max: float, device=None):
    """Corpus-level min/max over all ranks (the reference tracks these per feature while preprocessing)."""
    rank, world = rank_world()
    t = torch.tensor([-local_min, local_max], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(-t[0]), float(t[1])


def reduce_mean_scalar(x: torch.Tensor) -> torch.Tensor:
    """Mean of a scalar over ranks (equal per-rank batch sizes -> global mean of the loss)."""
    rank, world = rank_world()
    if world == 1:
        return x
    y = x.detach().clone()
    torch.distributed.all_reduce(y, op=torch.distributed.ReduceOp.SUM)
    return y / world


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin this process to the CPUs that are local to its GPU (NVML's ideal CPU affinity) BEFORE it allocates pinned host
    memory: first-touch then places the staging buffers on the GPU's NUMA node, so that the ranks of one box do not
    share one socket's memory controllers for their host<->device copies.  Returns False (and changes nothing) when
    NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return False
        os.sched_setaffinity(0, allowed)
        return True
    except Exception:
        return False
