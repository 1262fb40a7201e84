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
