"""The gateway's copy of the policy evaluator (gateway/policy_bundles.go:1132-1231 evaluatePolicyCheck), which
handleSimulatePolicyBundle (:322-368) runs against the published bundles with one bundle swapped for a draft, and the
request runPolicySimulation builds for a pack's policy tests (gateway/packs.go:1725-1771).  Same decisions as the safety
kernel; the effective-config reasons print the topic with %q instead of '%s' (:1207,:1211).

CPU: the two restatements against each other in that flavor.  GPU: the product (GatewayPolicyEvaluator over a scratch
engine, reasons through cordum_reason_flavor) against the oracle, strings included."""
import random
import sys

import pytest

import oracle_lib
from cordum_b200 import wire

sys.path.insert(0, oracle_lib.ORACLE_DIR)
import py_oracle  # noqa: E402

# gateway/policy_bundles_test.go:15-21
POLICY_CONTENT = {"rules": [{"id": "allow-all", "match": {"topics": ["job.*"]}, "decision": "allow"}]}

DRAFT = {"default_tenant": "default",
         "tenants": {"default": {"deny_topics": ["job.\"quoted\".*"], "mcp": {"deny_servers": ["Evil.Example"], "allow_actions": ["read"]}}},
         "rules": [{"id": "gate", "decision": "require_approval", "reason": "needs a human", "match": {"topics": ["job.prod.*"]}},
                   {"id": "cons", "decision": "allow", "match": {"topics": ["job.lim.*"]}, "constraints": {"budgets": {"max_retries": 2}}},
                   {"id": "allow-all", "match": {"topics": ["job.*"]}, "decision": "allow"}]}

TOPICS = ["job.test", "job.prod.deploy", "job.lim.a", 'job."quoted".x', "job.tab\there", "job.café", "job.nbsp x", "job.\U0001f600",
          "job.ctl\x01\x7f", "job.back\\slash", "", "  ", "notjob", " job.padded ", "job.zwsp​", "job.͸unassigned"]
EFFS = [None, b'{"safety":{"denied_topics":["job.*"]}}', b'{"safety":{"allowed_topics":["job.only"]}}',
        b'{"data":{"safety":{"denied_topics":["job.tab*","job.c*"],"allowed_topics":["job.*"]}}}',
        b'{"safety":{"mcp":{"deny_tools":["RM"]}}}', b"not json"]
LABELS = [None, {"mcp.server": " Evil.Example "}, {"mcp_server": "fine", "mcp.action": "WRITE"}, {"mcpTool": "rm", "mcp.server": "ok"},
          {"mcp.server": "cafÉ\t"}, {"mcp.tool": "a\"b\\c"}]


def requests(n=400, seed=5):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        r = {"job_id": "job-%d" % i, "topic": rng.choice(TOPICS)}
        if rng.random() < 0.7:
            r["tenant"] = rng.choice(["default", "other", ""])
        eff = rng.choice(EFFS)
        if eff is not None:
            r["effective_config"] = eff
        lab = rng.choice(LABELS)
        if lab is not None:
            r["labels"] = dict(lab)
        out.append(r)
    return out


def test_restatements_agree_in_the_gateway_flavor():
    o = oracle_lib.Oracle(DRAFT, {"topics": {}, "pools": {}}, [])
    seen = set()
    for req in requests():
        want = py_oracle.kernel_evaluate(DRAFT, req, gateway=True)
        got = o.eval_one(req, wire.MODE_POLICY_ONLY, flavor=1)
        assert got["decision"] == want["decision"] and got["reason"] == want["reason"] and got["rule_id"] == want["rule_id"], (req, got, want)
        kern = o.eval_one(req, wire.MODE_POLICY_ONLY, flavor=0)
        assert kern["decision"] == got["decision"] and kern["rule_idx"] == got["rule_idx"]      # only the wording may differ
        if kern["reason"] != got["reason"]:
            assert "effective config" in got["reason"]
            seen.add(got["reason"])
    assert 'topic "job.tab\\there" denied by effective config' in seen
    assert 'topic "job.\\"quoted\\".x" not allowed by effective config' in seen
    assert any("\\u00a0" in s for s in seen) and any("\\x01\\x7f" in s for s in seen) and any("café" in s for s in seen)
    o.close()


def test_pack_simulation_request_mapping():   # gateway/packs.go:1725-1760
    from cordum_b200 import reference_api as api

    with pytest.raises(ValueError):
        api.pack_simulation_request({"tenant_id": "t"}, "pack-1")
    r = api.pack_simulation_request({"topic": "job.x", "capability": "cap", "risk_tags": ["a"]}, "pack-1", default_tenant="acme")
    assert r == {"topic": "job.x", "tenant": "", "meta": {"tenant_id": "acme", "capability": "cap", "risk_tags": ["a"], "requires": [],
                                                          "pack_id": "pack-1", "actor_id": "", "actor_type": ""}}
    r = api.pack_simulation_request({"topic": "job.x", "tenant_id": "t1", "pack_id": "own", "actor_id": ""}, "pack-1",
                                    auth={"tenant": "t2", "principal_id": "alice"})
    assert r["tenant"] == "t2" and r["meta"]["tenant_id"] == "t2" and r["meta"]["actor_id"] == "alice" and r["meta"]["pack_id"] == "own"


@pytest.mark.gpu
def test_bundle_simulation_like_the_reference_test():   # gateway/policy_bundles_test.go:72-93
    from cordum_b200 import reference_api as api

    gw = api.GatewayPolicyEvaluator()
    resp = gw.evaluate_policy_check(POLICY_CONTENT, "cfg:abc", {"topic": "job.test", "tenant": "default"})
    assert resp["decision"] == "ALLOW" and resp["policy_snapshot"] == "cfg:abc" and resp["rule_id"] == "allow-all"
    assert gw.evaluate_policy_check(None, "", {"topic": "job.test"})["decision"] == "ALLOW"          # nil policy
    r = gw.evaluate_policy_check(None, "s", {"topic": "nope"})
    assert r["decision"] == "DENY" and r["reason"] == "unsupported topic" and r["policy_snapshot"] == ""


@pytest.mark.gpu
def test_gateway_evaluator_matches_the_oracle_strings_included():
    from cordum_b200 import reference_api as api

    gw = api.GatewayPolicyEvaluator()
    live = api.SafetyKernelServer()
    live.set_policy(POLICY_CONTENT, "live-1")
    o = oracle_lib.Oracle(DRAFT, {"topics": {}, "pools": {}}, [])
    reqs = requests()
    got = gw.evaluate_batch(DRAFT, "draft-7", reqs)
    n_q = 0
    for req, g in zip(reqs, got):
        w = o.eval_one(req, wire.MODE_POLICY_ONLY, flavor=1)
        assert (g["decision"], g["reason"], g["rule_id"], g["approval_required"]) == \
               (w["decision"], w["reason"], w["rule_id"], w["approval_required"]), (req, g, w)
        assert g["approval_ref"] == (req["job_id"] if w["approval_required"] else "")
        assert g["policy_snapshot"] == ("" if w["reason"] in ("missing topic", "unsupported topic") else "draft-7")
        n_q += '\\' in g["reason"]
    assert n_q > 20
    # the same requests through the safety-kernel flavour of the same engine: '%s' wording, request spelling in MCP reasons
    kern = gw._server.evaluate_batch(reqs)
    for req, g in zip(reqs, kern):
        w = o.eval_one(req, wire.MODE_POLICY_ONLY, flavor=0)
        assert (g["decision"], g["reason"]) == (w["decision"], w["reason"]), (req, g, w)
    # simulating against the draft left the live server's policy alone
    assert live.check({"topic": "job.prod.deploy"})["decision"] == "ALLOW" and live.engine.current_snapshot() == "live-1"
    o.close()
