"""Bounded dynamic dictionaries on the device path: when the topic / effective-config dictionary is full the engine starts
a new generation (host.cpp reset_dynamic_dictionaries) - tables and the device-side dictionaries are re-uploaded, batches
encoded under the old ids are refused as stale, strings of batches dispatched earlier stay those of their own generation."""
import numpy as np
import pytest

import kats
import oracle_lib
from cordum_b200 import wire

pytestmark = pytest.mark.gpu
FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")


@pytest.mark.parametrize("device_encode", [False, True], ids=["host-encode", "device-encode"])
def test_generations_on_the_gpu(device_encode):
    from cordum_b200 import engine
    from cordum_b200.engine import CordumError

    policy = {"rules": [{"id": "d", "decision": "deny", "reason": "no", "match": {"topics": ["job.bad.*"]}}]}
    routing = {"topics": {"job.keep.a": ["p"]}, "pools": {"p": {}}}
    workers = [kats.hb("", "p")]            # a worker without an id: the subject is the job's own topic (bus/nats.go:131-135)
    e = engine.Engine(device=0, max_topics=24, max_effcfgs=8)
    e.load_policy(policy, "gen")
    e.load_routing(routing)
    e.load_workers(workers)
    o = oracle_lib.Oracle(policy, routing, workers)
    b, old = e.batch(32), e.batch(32)
    first = [{"topic": "job.keep.a"}, {"topic": "job.first.one", "effective_config": b'{"safety":{"denied_topics":["job.first.*"]}}'}]
    (old.encode_device(first) if device_encode else old.encode(first)).dispatch()
    assert old.subject(0) == "job.keep.a" and old.reason(1) == "topic 'job.first.one' denied by effective config"
    held = e.batch(32)
    held.encode(first)                      # encoded, not dispatched: will be stale after a reset
    for rnd in range(8):
        jobs = [{"topic": "job.%s.r%d-%d" % ("bad" if i % 2 else "ok", rnd, i)} for i in range(9)] + [{"topic": "job.keep.a"}]
        got = (b.encode_device(jobs) if device_encode else b.encode(jobs)).dispatch()
        want = o.eval(jobs)
        for f in FIELDS:
            assert np.array_equal(got[f], want[f]), (rnd, f)
        assert b.subject(9) == "job.keep.a"
    with pytest.raises(CordumError):        # ids of a previous generation are never evaluated
        held.dispatch()
    # the batch dispatched before the resets still names its own topics
    assert old.subject(0) == "job.keep.a" and old.reason(1) == "topic 'job.first.one' denied by effective config"
    got = held.encode(first).dispatch()
    want = o.eval(first)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    e.close()
    o.close()
