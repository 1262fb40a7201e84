"""The reference-facing mirror (cordum_b200/reference_api.py): CPU tests for the pure mapping
logic, GPU tests that read like the reference's own strategy / kernel tests."""
import pytest

from cordum_b200 import reference_api as api
from cordum_b200 import wire


# ---------------------------------------------------------------- CPU: field mapping, registry
def test_extract_tenant():   # controlplane/scheduler/helpers_test.go:36-55
    assert api.extract_tenant(None) == api.DEFAULT_TENANT
    assert api.extract_tenant({"tenant_id": "t1", "env": {"tenant_id": "t2"}, "principal_id": "p1"}) == "t1"
    assert api.extract_tenant({"env": {"tenant_id": "t2"}}) == "t2"
    assert api.extract_tenant({"principal_id": "p2"}) == "default"


def test_policy_check_request_mapping():   # controlplane/scheduler/safety_client.go:80-95
    req = {"job_id": "j1", "topic": "job.x", "tenant_id": "t1", "principal_id": "p", "labels": {"a": "b"},
           "meta": {"capability": "c"}, "env": {api.EFFECTIVE_CONFIG_ENV: '{"safety":{}}', "tenant_id": "ignored"}}
    out = api.policy_check_request(req)
    assert out["tenant"] == "t1" and out["topic"] == "job.x" and out["job_id"] == "j1" and out["principal_id"] == "p"
    assert out["labels"] == {"a": "b"} and out["meta"] == {"capability": "c"}
    assert out["effective_config"] == b'{"safety":{}}'
    assert "effective_config" not in api.policy_check_request({"topic": "job.x", "env": {api.EFFECTIVE_CONFIG_ENV: ""}})


def test_decision_from_proto():   # controlplane/scheduler/safety_client_test.go:200-214
    assert api.decision_from_proto(wire.DEC_ALLOW) == "ALLOW"
    assert api.decision_from_proto(wire.DEC_DENY) == "DENY"
    assert api.decision_from_proto(wire.DEC_REQUIRE_HUMAN) == "REQUIRE_APPROVAL"
    assert api.decision_from_proto(wire.DEC_THROTTLE) == "THROTTLE"
    assert api.decision_from_proto(wire.DEC_ALLOW_WITH_CONSTRAINTS) == "ALLOW_WITH_CONSTRAINTS"
    assert api.decision_from_proto(wire.DEC_UNSPECIFIED) == "DENY"


def test_memory_registry():   # controlplane/scheduler/registry_memory_test.go:11-58 + TTL registry_memory.go:23,76-81
    now = [0.0]
    r = api.MemoryRegistry(clock=lambda: now[0])
    r.update_heartbeat({"worker_id": "worker-1", "pool": "gpu-pool", "cpu_load": 50.0})
    snap = r.snapshot()
    assert len(snap) == 1 and snap["worker-1"]["pool"] == "gpu-pool"
    r.update_heartbeat({"worker_id": "w1", "pool": "A"})
    r.update_heartbeat({"worker_id": "w2", "pool": "A"})
    r.update_heartbeat({"worker_id": "w3", "pool": "B"})
    assert len(r.workers_for_pool("A")) == 2 and len(r.workers_for_pool("B")) == 1 and len(r.workers_for_pool("C")) == 0
    r.update_heartbeat(None)
    r.update_heartbeat({"worker_id": ""})
    assert len(r.snapshot()) == 4
    now[0] = 29.0
    r.update_heartbeat({"worker_id": "w1", "pool": "A"})
    now[0] = 31.0                                  # everything but w1 is older than the 30 s TTL
    assert set(r.snapshot()) == {"w1"}
    r.expire()
    assert set(r._hb) == {"w1"}


# ---------------------------------------------------------------- GPU: the seams end to end
def routing_for_topic(topic, pool):
    return {"topics": {topic: [pool]}, "pools": {pool: {}}}


def hb(worker_id, pool, active=0, cpu=0.0, gpu=0.0, maxp=0, labels=None):
    return {"worker_id": worker_id, "pool": pool, "active_jobs": active, "max_parallel_jobs": maxp, "cpu_load": cpu,
            "gpu_utilization": gpu, "labels": labels or {}}


@pytest.fixture(scope="module")
def eng():
    from cordum_b200 import engine

    e = engine.Engine(device=0)
    yield e
    e.close()


@pytest.mark.gpu
def test_strategy_like_the_reference_tests(eng):   # controlplane/scheduler/strategy_least_loaded_test.go:20-144
    s = api.LeastLoadedStrategy(routing_for_topic("job.default", "default"), engine=eng)
    workers = {"w1": hb("w1", "default", 2, 50), "w2": hb("w2", "default", 1, 10), "w3": hb("w3", "other", 0, 0)}
    assert s.pick_subject({"topic": "job.default"}, workers) == "worker.w2.jobs"
    with pytest.raises(api.ErrNoWorkers):
        s.pick_subject({"topic": "job.default"}, {})
    with pytest.raises(api.ErrNoPoolMapping):
        s.pick_subject({"topic": "job.unknown"}, {})
    with pytest.raises(ValueError, match="missing topic"):
        s.pick_subject({"topic": ""}, workers)
    workers = {"w1": hb("w1", "default", 5, 90), "w2": hb("w2", "default", 2, 50), "w3": hb("w3", "default", 1, 10)}
    assert s.pick_subject({"topic": "job.default", "labels": {"preferred_worker_id": "w2"}}, workers) == "worker.w2.jobs"
    with pytest.raises(api.ErrPoolOverloaded):
        s.pick_subject({"topic": "job.default"}, {"w1": hb("w1", "default", 1, 1, maxp=1)})
    # only the loads moved: the registry map is re-used, deltas are applied
    workers = {"w1": hb("w1", "default", 0, 1, maxp=1)}
    assert s.pick_subject({"topic": "job.default"}, workers) == "worker.w1.jobs"
    workers["w1"]["active_jobs"] = 1
    with pytest.raises(api.ErrPoolOverloaded):
        s.pick_subject({"topic": "job.default"}, workers)
    # UpdateRouting / CurrentRouting (strategy_least_loaded.go:28-38)
    s.update_routing({"topics": {"job.other": "p2"}, "pools": {"p2": {"requires": ["gpu"]}}})
    assert s.current_routing() == {"topics": {"job.other": ["p2"]}, "pools": {"p2": {"requires": ["gpu"]}}}
    assert s.pick_subject({"topic": "job.other", "meta": {"requires": ["GPU"]}}, {"g": hb("g", "p2")}) == "worker.g.jobs"


@pytest.mark.gpu
def test_safety_kernel_surface(eng):   # controlplane/safetykernel/kernel_test.go:16-46,79-145,225-282
    srv = api.SafetyKernelServer(engine=eng)
    srv.set_policy({"default_tenant": "default",
                    "tenants": {"default": {"allow_topics": ["job.*"], "mcp": {"deny_servers": ["blocked.example.com"]}}}}, "snap-a")
    resp = srv.check({"job_id": "job-1", "topic": "job.default", "tenant": "default",
                      "labels": {"mcp.server": "blocked.example.com", "mcp.tool": "read"}})
    assert resp["decision"] == "DENY" and resp["reason"] == 'mcp server "blocked.example.com" denied'
    srv.set_policy({"default_tenant": "default", "rules": [
        {"id": "deny-delete", "decision": "deny", "match": {"tenants": ["default"], "topics": ["job.db.delete"]},
         "remediations": [{"id": "archive", "title": "Archive instead", "replacement_topic": "job.db.archive"}]},
        {"id": "gate", "decision": "require_approval", "reason": "needs a human", "match": {"topics": ["job.prod.*"]},
         "constraints": {"budgets": {"max_runtime_ms": 1000}}}]}, "snap-b")
    resp = srv.check({"job_id": "job-5", "topic": "job.db.delete", "tenant": "default"})
    assert resp["decision"] == "DENY" and len(resp["remediations"]) == 1
    assert resp["remediations"][0]["replacement_topic"] == "job.db.archive" and resp["policy_snapshot"] == "snap-b"
    for call in (srv.evaluate, srv.explain, srv.simulate):
        r = call({"job_id": "job-9", "topic": "job.prod.deploy", "tenant": "default"})
        assert r["decision"] == "REQUIRE_HUMAN" and r["approval_required"] and r["approval_ref"] == "job-9"
        assert r["reason"] == "needs a human" and r["constraints"]["budgets"]["max_runtime_ms"] == 1000
    r = srv.evaluate({})
    assert r["decision"] == "DENY" and r["reason"] == "missing topic" and r["policy_snapshot"] == ""
    r = srv.check({"job_id": "job-3", "topic": "job.deny", "tenant": "default",
                   "effective_config": b'{"safety":{"denied_topics":["job.deny"]}}'})
    assert r["decision"] == "DENY" and "denied" in r["reason"]
    # ListSnapshots keeps 10, newest first (kernel_test.go:252-268)
    for i in range(12):
        srv.set_policy(None, "snap-%d" % i)
    snaps = srv.list_snapshots()
    assert len(snaps) == 10 and snaps[0] == "snap-11"
    assert srv.check({"topic": "job.anything"})["decision"] == "ALLOW"


@pytest.mark.gpu
def test_scheduler_decision_switch(eng):   # engine.go:294-347,393; engine_test.go:332-356; integration_test.go:48-135
    sch = api.Scheduler(engine=eng)
    sch.kernel.set_policy({"default_tenant": "default", "rules": [
        {"id": "gate", "decision": "require_approval", "reason": "prod", "match": {"topics": ["job.prod.*"]}},
        {"id": "nosys", "decision": "deny", "reason": "no", "match": {"topics": ["job.sys.*"]}}]}, "s")
    sch.strategy.update_routing({"topics": {"job.default": ["default"], "job.prod.x": ["default"]}, "pools": {"default": {}}})
    workers = {"w1": hb("w1", "default", 0, 5, maxp=4)}
    jobs = [{"job_id": "a", "topic": "job.default", "tenant_id": "default"},
            {"job_id": "b", "topic": "job.prod.x", "env": {"tenant_id": "default"}},
            {"job_id": "c", "topic": "job.sys.destroy"},
            {"job_id": "d", "topic": "sys.destroy"},
            {"job_id": "e", "topic": "job.unmapped"}]
    out = sch.process_jobs(jobs, workers)
    assert out[0]["decision"] == "ALLOW" and out[0]["subject"] == "worker.w1.jobs"
    assert out[1]["decision"] == "REQUIRE_APPROVAL" and out[1]["subject"] == "" and out[1]["approval_required"]
    assert out[2]["decision"] == "DENY" and out[2]["reason"] == "no" and out[2]["subject"] == ""
    assert out[3]["decision"] == "DENY" and out[3]["reason"] == "unsupported topic"
    assert out[4]["decision"] == "ALLOW" and isinstance(out[4]["error"], api.ErrNoPoolMapping)
    # the approved replay goes straight to routing (engine.go:484-522)
    out = sch.process_jobs([jobs[1]], workers, approved=[True])
    assert out[0]["decision"] == "ALLOW" and out[0]["subject"] == "worker.w1.jobs" and out[0]["reason"] == "approval granted"
