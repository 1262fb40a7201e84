"""The device-side encoder (cordum_b200/csrc/encode.cu) against its specification, the host encoder (host.cpp
Host::encode_job): same records for every job, and the same decisions.  Run on the B200 box: pytest -m gpu."""
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


def by_job(batch):
    job, route, slot = batch.records()
    assert np.array_equal(job["orig"][slot], np.arange(len(job), dtype=np.uint32)), "slot_of / orig are not inverse"
    assert np.all(np.diff(job["topic"].astype(np.int64)) >= 0), "records are not grouped by topic"
    return job[slot], route[slot]


def assert_records_equal(a, b, what):
    ja, ra = a
    jb, rb = b
    for f in ja.dtype.names:
        if f == "spare":
            continue
        bad = np.nonzero((ja[f] != jb[f]).reshape(len(ja), -1).any(axis=1))[0]
        assert len(bad) == 0, "%s: job record field %s differs at jobs %s: %s vs %s" % (what, f, bad[:6], ja[f][bad[:6]], jb[f][bad[:6]])
    for f in ra.dtype.names:
        bad = np.nonzero(ra[f] != rb[f])[0]
        assert len(bad) == 0, "%s: route record field %s differs at jobs %s: %s vs %s" % (what, f, bad[:6], ra[f][bad[:6]], rb[f][bad[:6]])


@pytest.mark.parametrize("name,n", [("tiny", 2000), ("c2", None), ("c3", 300_000)])
def test_device_records_equal_host_records(eng, name, n):
    cfg = synth.make_config(name, n)
    eng.load_policy(cfg.policy, "t")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    bh, bd = eng.batch(cfg.jobs.n_jobs), eng.batch(cfg.jobs.n_jobs)
    bh.encode(cfg.jobs)                                # registers every topic / effective config of the batch ...
    host = by_job(bh.encode(cfg.jobs))                 # ... so this encode sees the final dictionaries (an effective config
    #                                                    seen for the first time interns its MCP values: a job encoded
    #                                                    before that carries the equivalent id "other" for them)
    f0 = eng.host_fallbacks()
    for env in (cfg.jobs, cfg.jobs.deinterned(), eng.pinned_envelopes(cfg.jobs.deinterned())):
        dev = by_job(bd.encode_device(env))
        assert_records_equal(dev, host, "%s (%s)" % (name, type(env).__name__))
    assert eng.host_fallbacks() == f0, "the device encoder should not have needed the host for a known vocabulary"
    want = bh.dispatch().copy()
    got = bd.dispatch()
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    bh.free()
    bd.free()


def test_first_sight_and_non_ascii_fall_back_to_the_host(eng):
    """Unknown topics / effective configs and non-ASCII text are the host's: the batch is re-encoded there at wait time
    and the decisions are the oracle's either way."""
    policy = {"default_tenant": "default", "rules": [
        {"id": "deny-kelvin", "decision": "deny", "reason": "no", "match": {"tenants": ["KELVIN"]}},
        {"id": "deny-a", "decision": "deny", "reason": "x", "match": {"topics": ["job.a.*"], "capabilities": ["cap.one"]}}]}
    routing = {"topics": {"job.b.one": ["p"]}, "pools": {"p": {}}}
    workers = [{"worker_id": "w", "pool": "p"}]
    eng.load_policy(policy, "t")
    eng.load_routing(routing)
    eng.load_workers(workers)
    o = oracle_lib.Oracle(policy, routing, workers)
    jobs = [{"topic": "job.a.new-%d" % (i % 7), "meta": {"capability": " CAP.ONE " if i % 2 else "cap.two"}} for i in range(64)]
    jobs += [{"topic": "job.b.one", "tenant": "Kelvin"}, {"topic": "job.b.one", "tenant": "kelvin"},
             {"topic": "job.b.one", "effective_config": b'{"safety":{"denied_topics":["job.b.*"]}}'},
             {"topic": "job.b.one", "labels": {"mcp.server": "sérver"}}, {"topic": " job.b.one"}]
    b = eng.batch(len(jobs))
    f0 = eng.host_fallbacks()
    got = b.encode_device(jobs).dispatch().copy()
    assert eng.host_fallbacks() == f0 + 1
    want = o.eval(jobs)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    assert b.reason(len(jobs) - 3) == "topic 'job.b.one' denied by effective config"
    # the ASCII part of the vocabulary is known now: the same jobs without the non-ASCII ones stay on the device
    ascii_jobs = [j for j in jobs if "Kelvin" not in str(j) and "sérver" not in str(j)]
    f1 = eng.host_fallbacks()
    got = b.encode_device(ascii_jobs).dispatch().copy()
    assert eng.host_fallbacks() == f1
    want = o.eval(ascii_jobs)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    assert b.reason(len(ascii_jobs) - 2) == "topic 'job.b.one' denied by effective config"   # host copy of the records fetched lazily
    b.free()


def test_ragged_and_empty_device_batches(eng):
    cfg = synth.make_config("tiny", 300)
    eng.load_policy(cfg.policy, "t")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    want = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers).eval(cfg.jobs)
    b = eng.batch(300)
    b.encode(cfg.jobs).dispatch()   # vocabulary
    for n in (1, 2, 31, 32, 33, 255, 300):
        got = b.encode_device(cfg.jobs.slice(0, n)).dispatch()
        for f in FIELDS:
            assert np.array_equal(got[f], want[f][:n]), (n, f)
    b.free()
