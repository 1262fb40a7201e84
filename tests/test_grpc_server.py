"""The gRPC SafetyKernel service (cordum_b200/grpc_server.py): Check / Evaluate / Explain / Simulate / ListSnapshots over a
real grpc server on localhost (kernel.go:106-127).  CPU: the wire layer alone, with the Python oracle standing in as the
evaluator (test infrastructure).  GPU: the engine behind it, responses compared with the oracle's."""
import sys
import threading

import grpc
import pytest

import oracle_lib
from cordum_b200 import grpc_server as gs

sys.path.insert(0, oracle_lib.ORACLE_DIR)
import py_oracle  # noqa: E402

POLICY = {"default_tenant": "default",
          "tenants": {"default": {"mcp": {"deny_servers": ["blocked.example.com"]}}},
          "rules": [
              {"id": "deny-delete", "decision": "deny", "reason": "no deletes", "match": {"tenants": ["default"], "topics": ["job.db.delete"]},
               "remediations": [{"id": "archive", "title": "Archive instead", "replacement_topic": "job.db.archive", "add_labels": {"mode": "archive"}}]},
              {"id": "gate", "decision": "require_approval", "reason": "needs a human", "match": {"topics": ["job.prod.*"]},
               "constraints": {"budgets": {"max_runtime_ms": 1000, "max_retries": 2}, "redaction_level": "strict"}},
              {"id": "svc-only", "decision": "deny", "reason": "humans only", "match": {"topics": ["job.hr.*"], "actor_types": ["service"]}}]}
REQUESTS = [
    {"job_id": "j1", "topic": "job.db.delete", "tenant": "default"},
    {"job_id": "j2", "topic": "job.prod.deploy", "tenant": "default", "meta": {"actor_id": "alice", "actor_type": "human", "risk_tags": ["write"]}},
    {"job_id": "j3", "topic": "job.default", "tenant": "default", "labels": {"mcp.server": "Blocked.Example.com", "mcp.tool": "read"}},
    {"job_id": "j4", "topic": "job.hr.read", "meta": {"actor_type": "service"}},
    {"job_id": "j5", "topic": "job.hr.read", "principal_id": "svc"},                  # nil meta: only the principal is looked at
    {"job_id": "j6", "topic": "job.x", "effective_config": b'{"safety":{"denied_topics":["job.x"]}}'},
    {"job_id": "j7", "topic": ""},
    {"job_id": "j8", "topic": "nope"},
    {"job_id": "j9", "topic": "job.café", "tenant": "default", "labels": {"k": "v☃"}},
]


def oracle_response(req: dict, snapshot="snap-1") -> dict:
    r = py_oracle.kernel_evaluate(POLICY, req)
    rule = POLICY["rules"][r["rule_idx"]] if r["rule_idx"] >= 0 else {}
    return {"decision": r["decision"], "reason": r["reason"], "policy_snapshot": snapshot if r["has_snapshot"] else "",
            "rule_id": r["rule_id"], "constraints": rule.get("constraints") if r["has_constraints"] else None,
            "approval_required": r["approval_required"], "approval_ref": req.get("job_id", "") if r["approval_required"] else "",
            "remediations": rule.get("remediations", []) if r["has_snapshot"] else []}


def comparable(d: dict) -> dict:
    out = {k: d[k] for k in ("decision", "reason", "policy_snapshot", "rule_id", "approval_required", "approval_ref")}
    c = d.get("constraints")
    out["max_runtime_ms"] = (c or {}).get("budgets", {}).get("max_runtime_ms", 0) if c else None
    out["remediations"] = [(r["id"], r.get("replacement_topic", "")) for r in d.get("remediations") or []]
    return out


def run_against(servicer, want_of):
    server, port = gs.serve(servicer)
    try:
        with grpc.insecure_channel("127.0.0.1:%d" % port) as ch:
            stub = gs.SafetyKernelStub(ch)
            for req in REQUESTS:
                for call in (stub.Check, stub.Evaluate, stub.Explain, stub.Simulate):
                    got = gs.response_to_dict(call(gs.request_from_dict(req), timeout=20))
                    assert comparable(got) == comparable(want_of(req)), req
            snaps = list(stub.ListSnapshots(gs.ListSnapshotsRequest(), timeout=20).snapshots)
            # concurrent callers, one blocking RPC each
            errs = []

            def client(k):
                try:
                    for i in range(20):
                        req = REQUESTS[(k + i) % len(REQUESTS)]
                        got = gs.response_to_dict(stub.Check(gs.request_from_dict(req), timeout=20))
                        if comparable(got) != comparable(want_of(req)):
                            errs.append((req, got))
                except Exception as exc:   # noqa: BLE001
                    errs.append(exc)
            ts = [threading.Thread(target=client, args=(k,)) for k in range(16)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            assert not errs, errs[:2]
            return snaps
    finally:
        server.stop(0)


def test_wire_layer_round_trips_every_field():
    req = gs.request_from_dict(REQUESTS[1])
    back = gs.request_to_dict(gs.PolicyCheckRequest.FromString(req.SerializeToString()))
    assert back["meta"]["actor_type"] == 1 and back["meta"]["risk_tags"] == ["write"] and back["topic"] == "job.prod.deploy"
    assert "meta" not in gs.request_to_dict(gs.request_from_dict(REQUESTS[4]))          # nil Meta stays nil (kernel.go:349-356)
    assert "meta" in gs.request_to_dict(gs.request_from_dict({"topic": "job.x", "meta": {}}))
    r = gs.response_from_dict(oracle_response(REQUESTS[1]))
    assert r.constraints.budgets.max_retries == 2 and r.constraints.redaction_level == "strict" and r.approval_ref == "j2"
    r = gs.response_from_dict(oracle_response(REQUESTS[0]))
    assert r.remediations[0].add_labels["mode"] == "archive" and not r.HasField("constraints")


def test_service_over_localhost_with_the_python_oracle_as_evaluator():
    snaps = run_against(gs.SafetyKernelServicer(lambda req: oracle_response(req), lambda: ["snap-1", "snap-0"]), oracle_response)
    assert snaps == ["snap-1", "snap-0"]


def test_engine_failure_fails_closed_over_the_wire():
    def boom(req):
        raise RuntimeError("device lost")
    server, port = gs.serve(gs.SafetyKernelServicer(boom, lambda: []))
    try:
        with grpc.insecure_channel("127.0.0.1:%d" % port) as ch:
            resp = gs.SafetyKernelStub(ch).Check(gs.request_from_dict(REQUESTS[0]), timeout=20)
            assert gs.DECISION_NAME[resp.decision] == "DENY" and "safety kernel error: device lost" in resp.reason   # safety_client.go:98-101
    finally:
        server.stop(0)


@pytest.mark.gpu
def test_service_over_localhost_with_the_engine_behind_it():
    from cordum_b200 import engine

    eng = engine.Engine(device=0)
    eng.load_policy(POLICY, "snap-0")
    eng.load_policy(POLICY, "snap-1")
    eng.load_routing({"topics": {}, "pools": {}})
    eng.load_workers([])
    sv = gs.engine_servicer(eng, max_batch=64, max_wait_us=100)
    snaps = run_against(sv, oracle_response)
    assert snaps[:2] == ["snap-1", "snap-0"]
    st = sv.frontend.stats()
    assert st["requests"] >= len(REQUESTS) * 4 + 320 and st["batches"] <= st["requests"]     # concurrent RPCs share batches when they overlap
    sv.frontend.close()
    eng.close()
