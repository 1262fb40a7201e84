"""Job-batch data parallelism across the GPUs of one box.

Jobs are independent given (rule tables, routing tables, worker table) — the scheduler never
mutates worker load when it dispatches — so the batch shards by job index with the
tables replicated.  The one exchange step: each rank ingests the heartbeat deltas of its slice
of the worker registry and a single all-gather of the 16 B/worker load records rebuilds the
full load table on every rank before the per-pool argmin.
"""
from __future__ import annotations

import numpy as np


def job_range(rank: int, world: int, n_jobs: int) -> tuple[int, int]:
    """Contiguous, balanced-to-within-one partition of [0, n_jobs)."""
    per = (n_jobs + world - 1) // world
    return min(n_jobs, rank * per), min(n_jobs, (rank + 1) * per)


def worker_range(rank: int, world: int, n_workers: int) -> tuple[int, int]:
    """Slice of the worker registry (by slot) whose heartbeats this rank ingests.
    all_gather_into_tensor needs equal slices: the registry is padded to a multiple of `world`."""
    per = (n_workers + world - 1) // world
    return rank * per, (rank + 1) * per


def padded_workers(world: int, n_workers: int) -> int:
    return ((n_workers + world - 1) // world) * world


def gather_loads(local_slice, out=None):
    """All-gather of the per-rank load slices ((w1-w0) x 16 bytes, uint8 tensors) into the full
    slot-ordered table.  Works on any torch.distributed backend (nccl on the GPUs, gloo in tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_slice
    if out is None:
        out = torch.empty((local_slice.shape[0] * world,) + tuple(local_slice.shape[1:]), dtype=local_slice.dtype,
                          device=local_slice.device)
    dist.all_gather_into_tensor(out.view(-1), local_slice.contiguous().view(-1))
    return out


def loads_to_bytes(loads: np.ndarray) -> np.ndarray:
    """LOAD_DTYPE records -> (n, 16) uint8 view (what travels through the all-gather)."""
    return np.ascontiguousarray(loads).view(np.uint8).reshape(-1, 16)
