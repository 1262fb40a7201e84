"""Scheduler ticks (cordum_tick_async): one CUDA graph per tick = heartbeat epoch + policy of the new batch + route of the
previous batch.  Every batch must come out exactly as the oracle routes it on the loads of its own tick."""
import numpy as np
import pytest

import oracle_lib
from cordum_b200 import synth, wire

pytestmark = pytest.mark.gpu
FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")


@pytest.fixture(scope="module")
def eng():
    from cordum_b200 import engine

    e = engine.Engine(device=0)
    yield e
    e.close()


@pytest.mark.parametrize("name,n", [("tiny", 2000), ("c2", None)])
def test_ticks_match_oracle_epoch_by_epoch(eng, name, n):
    cfg = synth.make_config(name, n)
    eng.load_policy(cfg.policy, "t")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    W = cfg.workers.n_workers
    halves = [cfg.jobs.slice(0, cfg.jobs.n_jobs // 2), cfg.jobs.slice(cfg.jobs.n_jobs // 2, cfg.jobs.n_jobs - cfg.jobs.n_jobs // 2), cfg.jobs]
    batches = [eng.batch(h.n_jobs) for h in halves]
    for b, h in zip(batches, halves):
        b.encode(h).dispatch()          # resident
    rng = np.random.default_rng(5)
    loads_of, keep = {}, []
    for k in range(9):
        loads = cfg.workers.loads()
        loads["active_jobs"] = rng.integers(0, 9, W)
        loads["cpu_load"] = (rng.random(W) * 100).astype(np.float32)
        loads["gpu_utilization"] = (rng.random(W) * 100).astype(np.float32)
        keep.append(loads)
        i = k % 3
        batches[i].tick(loads.ctypes.data, 0, W)
        loads_of[i] = loads
        if k % 4 == 3:                    # now and then look at a batch in the middle of the pipeline
            j = (k - 1) % 3               # routed by the tick just issued
            batches[j].wait()
            o.update_workers(np.arange(W, dtype=np.uint32), loads_of[j])
            want = o.eval(halves[j])
            got = batches[j].fetch()
            for f in FIELDS:
                assert np.array_equal(got[f], want[f]), (k, j, f)
    for i in (2, 0, 1):                   # the last batch needs the flush (wait does it)
        batches[i].wait()
        o.update_workers(np.arange(W, dtype=np.uint32), loads_of[i])
        want = o.eval(halves[i])
        got = batches[i].fetch()
        for f in FIELDS:
            assert np.array_equal(got[f], want[f]), ("final", i, f)
    # a plain dispatch after ticks continues from the tables of the last tick
    last = loads_of[(9 - 1) % 3]
    o.update_workers(np.arange(W, dtype=np.uint32), last)
    got = batches[2].dispatch()
    want = o.eval(halves[2])
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), ("plain after ticks", f)
    # and ticks again after a plain dispatch
    batches[0].tick(keep[0].ctypes.data, 0, W)
    batches[0].wait()
    o.update_workers(np.arange(W, dtype=np.uint32), keep[0])
    got, want = batches[0].fetch(), o.eval(halves[0])
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), ("tick after plain", f)
    for b in batches:
        b.free()


def test_tick_argument_checks(eng):
    cfg = synth.make_config("tiny", 100)
    eng.load_policy(cfg.policy, "t")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    W = cfg.workers.n_workers
    loads = cfg.workers.loads()
    b = eng.batch(100)
    with pytest.raises(Exception):
        b.tick(loads.ctypes.data, 0, W)       # not encoded
    b.encode(cfg.jobs).dispatch()
    with pytest.raises(Exception):
        b.tick(loads.ctypes.data, 0, W - 1)   # not the whole table
    b.tick(loads.ctypes.data, 0, W)
    b.wait()
    b.free()
