"""world_size-2 gloo test of the multi-GPU path's host logic (no GPU): job sharding by index plus
one all-gather of per-rank worker-load slices must reproduce the single-process decisions.
The evaluator here is the oracle (this is a tests/ file); on the GPUs it is the CUDA engine."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from cordum_b200 import shard, synth, wire

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.make_config("tiny", 1001)          # odd size: ragged last shard
    W = cfg.workers.n_workers
    rng = np.random.default_rng(77)                # same stream on every rank
    new = cfg.workers.loads()
    new["active_jobs"] = rng.integers(0, 9, W)
    new["cpu_load"] = (rng.random(W) * 100).astype(np.float32)
    # each rank only knows the heartbeats of its own slice of the registry
    w0, w1 = shard.worker_range(rank, world, W)
    assert shard.padded_workers(world, W) == W
    mine = torch.from_numpy(shard.loads_to_bytes(new[w0:w1]).copy())
    full = shard.gather_loads(mine).numpy().view(wire.LOAD_DTYPE).reshape(-1)
    assert np.array_equal(full, new), "all-gathered load table differs from the global one"
    # evaluate this rank's job shard against the gathered table
    j0, j1 = shard.job_range(rank, world, cfg.jobs.n_jobs)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    o.update_workers(np.arange(W, dtype=np.uint32), full)
    part = o.eval(cfg.jobs, wire.MODE_POLICY_AND_ROUTE, first=j0, count=j1 - j0)
    np.save(os.path.join(tmp, "part%d.npy" % rank), part)
    dist.barrier()
    if rank == 0:
        parts = [np.load(os.path.join(tmp, "part%d.npy" % r)) for r in range(world)]
        whole = np.concatenate(parts)
        ref = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
        ref.update_workers(np.arange(W, dtype=np.uint32), new)
        want = ref.eval(cfg.jobs, wire.MODE_POLICY_AND_ROUTE)
        assert len(whole) == len(want)
        for f in ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot"):
            assert np.array_equal(whole[f], want[f]), f
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_ranges_partition_exactly():
    sys.path.insert(0, ROOT)
    from cordum_b200 import shard

    for n in (0, 1, 7, 8, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                a, b = shard.job_range(r, world, n)
                assert 0 <= a <= b <= n
                covered += list(range(a, b)) if n <= 1000 else []
            if n <= 1000:
                assert covered == list(range(n))
            assert shard.job_range(world - 1, world, n)[1] == n
