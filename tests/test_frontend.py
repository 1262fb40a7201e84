"""The micro-batching front-end (cordum_frontend_*): many threads each submitting ONE request per blocking call, as the
scheduler's processJob / the SafetyKernel RPC handlers do (engine.go:203-443; kernel.go:106-127).  Every response must be
the oracle's answer for that request alone, and the requests must actually have been served in batches."""
import threading

import numpy as np
import pytest

import kats
import oracle_lib
from cordum_b200 import frontend, synth, wire

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cordum_b200 import engine

    e = engine.Engine(device=0)
    yield e
    e.close()


def test_concurrent_single_requests_are_batched_and_exact(eng):
    cfg = synth.make_config("tiny", 1600)
    eng.load_policy(cfg.policy, "snap-fe")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    want = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers).eval(cfg.jobs)
    jobs = cfg.jobs.to_jobs()
    fe = frontend.Frontend(eng, max_batch=64, max_wait_us=500)
    n_threads, per = 32, len(jobs) // 32
    out = [None] * len(jobs)

    def client(t):
        for j in range(t * per, (t + 1) * per):
            out[j] = fe.submit(jobs[j])

    th = [threading.Thread(target=client, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    st = fe.stats()
    for j, r in enumerate(out):
        assert r.status == 0
        for f in ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot"):
            assert getattr(r.rec, f) == want[f][j], (j, f)
        routed = r.rec.route_status in (wire.ROUTE_OK, wire.ROUTE_OK_PREFERRED)
        assert (r.subject.decode() == "worker.%s.jobs" % cfg.workers.worker_id(int(r.rec.worker_slot))) if routed else r.subject == b""
        if r.rec.flags & wire.F_HAS_SNAPSHOT:
            assert r.snapshot == b"snap-fe"
            assert r.rule_id.decode() == (cfg.policy["rules"][r.rec.rule_idx]["id"] if r.rec.rule_idx >= 0 else "")
    assert st["requests"] == len(jobs) and st["batches"] < len(jobs) // 4, st   # served in batches, not one by one
    fe.close()


@pytest.mark.parametrize("case", kats.CASES[:40], ids=[c["name"] for c in kats.CASES[:40]])
def test_reference_kats_through_the_frontend(eng, case):
    """The reference's own known answers, one request per call (reason / rule id / subject strings included)."""
    eng.load_policy(case["policy"], "test")
    eng.load_routing(case["routing"])
    eng.load_workers(case["workers"])
    fe = frontend.Frontend(eng, max_batch=8, max_wait_us=0, mode=case["mode"])
    r = fe.submit(case["job"])
    fe.close()
    assert r.status == 0
    got = {"decision": wire.DEC_NAMES[r.rec.decision], "sched_decision": wire.DEC_NAMES[r.rec.sched_decision],
           "reason": r.reason.decode(), "rule_id": r.rule_id.decode(), "rule_idx": int(r.rec.rule_idx),
           "approval_required": bool(r.rec.flags & wire.F_APPROVAL_REQUIRED), "has_snapshot": bool(r.rec.flags & wire.F_HAS_SNAPSHOT),
           "has_constraints": bool(r.rec.flags & wire.F_CONSTRAINTS), "route": kats.ROUTE_NAMES[int(r.rec.route_status)],
           "subject": r.subject.decode(), "worker_slot": int(r.rec.worker_slot), "tie": bool(r.rec.flags & wire.F_TIE)}
    kats.check(case, got)


def test_oversized_request_fails_closed_alone(eng):
    cfg = synth.make_config("tiny", 10)
    eng.load_policy(cfg.policy, "t")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    fe = frontend.Frontend(eng, max_batch=4, max_wait_us=0, arena_bytes_per_request=256)
    big = {"topic": "job." + "x" * 5000}
    r = fe.submit(big)
    assert r.status != 0 and r.rec.decision == wire.DEC_DENY and b"safety kernel error" in r.reason
    ok = fe.submit(cfg.jobs.to_jobs(0, 1)[0])
    assert ok.status == 0
    fe.close()


def test_mcp_reason_quotes_the_requests_own_spelling(eng):
    """safety_policy.go:410,413 print %q of the label value as the request wrote it (trimmed; the action lower-cased,
    kernel.go:403); the job record only keeps the id of its case-folded form."""
    policy = {"default_tenant": "default", "tenants": {"default": {"mcp": {"deny_servers": ["evil.example"], "allow_actions": ["read"]}}}}
    eng.load_policy(policy, "t")
    eng.load_routing({"topics": {}, "pools": {}})
    eng.load_workers([])
    fe = frontend.Frontend(eng, max_batch=4, max_wait_us=0, mode=wire.MODE_POLICY_ONLY)
    o = oracle_lib.Oracle(policy, {"topics": {}, "pools": {}}, [])
    for labels in ({"mcp.server": "  EVIL.Example\t"}, {"mcp.server": "ok", "mcpAction": "WRITE \"x\""}, {"mcp_server": "Evil.EXAMPLE"}):
        job = {"topic": "job.x", "tenant": "default", "labels": labels}
        r = fe.submit(job)
        want = o.eval_one(job, wire.MODE_POLICY_ONLY)
        assert r.status == 0 and wire.DEC_NAMES[r.rec.decision] == want["decision"] == "DENY"
        assert r.reason.decode() == want["reason"]
    assert fe.submit({"topic": "job.x", "labels": {"mcp.server": "  EVIL.Example\t"}}).reason == b'mcp server "EVIL.Example" denied'
    fe.close()
    o.close()


def test_decision_cache_like_the_safety_kernels(eng):
    """SAFETY_DECISION_CACHE_TTL (kernel.go:149-162,250-303): a request identical to one evaluated under the same policy
    within the TTL is answered from memory; the early topic denials are never stored; a new policy is a new key space."""
    import time

    policy = {"default_tenant": "default", "rules": [
        {"id": "gate", "decision": "require_approval", "reason": "needs a human", "match": {"topics": ["job.prod.*"]}},
        {"id": "lab", "decision": "deny", "reason": "no", "match": {"topics": ["job.l.*"], "labels": {"a": "1", "b": "2"}}}]}
    eng.load_policy(policy, "snap-c1")
    eng.load_routing({"topics": {}, "pools": {}})
    eng.load_workers([])
    fe = frontend.Frontend(eng, max_batch=8, max_wait_us=0, mode=wire.MODE_POLICY_ONLY, cache_ttl_us=200_000)
    req = {"topic": "job.prod.deploy", "tenant": "default", "meta": {"risk_tags": ["x", "y"]}}
    r1 = fe.submit(req)
    r2 = fe.submit(dict(req))
    assert fe.cache_stats() == {"hits": 1, "misses": 1, "entries": 1}
    assert bytes(r1) == bytes(r2) and r2.rule_id == b"gate" and r2.snapshot == b"snap-c1"
    assert r2.rec.flags & wire.F_APPROVAL_REQUIRED                         # the adapter derives approval_ref from this, per request
    # a map has no order; a repeated field has
    a = fe.submit({"topic": "job.l.x", "labels": {"a": "1", "b": "2"}})
    b = fe.submit({"topic": "job.l.x", "labels": {"b": "2", "a": "1"}})
    assert a.rec.decision == b.rec.decision == wire.DEC_DENY and fe.cache_stats()["hits"] == 2
    fe.submit({"topic": "job.prod.deploy", "tenant": "default", "meta": {"risk_tags": ["y", "x"]}})
    assert fe.cache_stats()["hits"] == 2 and fe.cache_stats()["entries"] == 3
    # ("ab","c") vs ("a","bc"): lengths are part of the key
    fe.submit({"topic": "job.k", "tenant": "ab", "principal_id": "c"})
    fe.submit({"topic": "job.k", "tenant": "a", "principal_id": "bc"})
    assert fe.cache_stats()["hits"] == 2
    # early denials are returned before the cache is written (kernel.go:171-176 vs :250)
    for _ in range(2):
        assert fe.submit({"topic": "nope"}).reason == b"unsupported topic"
    assert fe.cache_stats()["hits"] == 2
    # a reload: new generation, nothing of the old policy is served
    eng.load_policy({"rules": []}, "snap-c2")
    r3 = fe.submit(req)
    assert r3.rec.decision == wire.DEC_ALLOW and r3.snapshot == b"snap-c2" and fe.cache_stats()["hits"] == 2
    assert fe.submit(req).rec.decision == wire.DEC_ALLOW and fe.cache_stats()["hits"] == 3
    # expiry
    time.sleep(0.25)
    n = fe.cache_stats()["hits"]
    fe.submit(req)
    assert fe.cache_stats()["hits"] == n
    fe.close()
    # off by default, and never for front-ends that route (the answer depends on the heartbeats)
    for kw in ({"mode": wire.MODE_POLICY_ONLY}, {"mode": wire.MODE_POLICY_AND_ROUTE, "cache_ttl_us": 10**6}):
        f2 = frontend.Frontend(eng, max_batch=8, max_wait_us=0, **kw)
        f2.submit(req); f2.submit(req)
        assert f2.cache_stats()["hits"] == 0
        f2.close()
