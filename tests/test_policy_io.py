"""Host-side document loading (cordum_b200/policy_io.py) against the reference's own config tests:
infra/config/pools_test.go:9-45, safety_policy_test.go:5-13,143-151, and kernel.go:694-778 (mergePolicies)."""
import pytest

from cordum_b200 import policy_io


def test_load_pool_config_success():   # pools_test.go:9-30
    body = b"topics:\n  job.default: default\n  job.batch:\n    - batch\n    - batch-b\npools:\n  default:\n    requires: [\"docker\", \"git\"]\n"
    cfg = policy_io.parse_pools_config(body)
    assert cfg["topics"]["job.default"] == ["default"]
    assert cfg["topics"]["job.batch"] == ["batch", "batch-b"]
    assert cfg["pools"]["default"]["requires"][0] == "docker"


def test_load_pool_config_errors():   # pools_test.go:32-45; pools.go parseTopicPools
    with pytest.raises(ValueError, match="no topics"):
        policy_io.parse_pools_config(b"topics: {}\n")
    for bad in (b"topics:\n  job.a: ''\n", b"topics:\n  job.a: []\n", b"topics:\n  job.a: [x, '']\n", b"topics:\n  job.a: 7\n"):
        with pytest.raises(ValueError):
            policy_io.parse_pools_config(bad)


def test_parse_safety_policy_empty_and_invalid_decision():   # safety_policy_test.go:5-13,143-151
    assert policy_io.parse_safety_policy(b"") is None and policy_io.parse_safety_policy(None) is None
    with pytest.raises(ValueError, match="invalid decision"):
        policy_io.parse_safety_policy("rules:\n  - id: r\n    decision: maybe\n")
    p = policy_io.parse_safety_policy("rules:\n  - id: r\n    decision: allow\n")
    assert p["rules"][0]["decision"] == "allow" and p["tenants"] == {}


def test_merge_policies_semantics():   # kernel.go:694-778
    base = policy_io.parse_safety_policy("version: v1\ntenants:\n  t:\n    allow_topics: [job.a]\n    max_concurrent_jobs: 5\n    mcp:\n      deny_tools: [rm]\nrules:\n  - id: b\n    decision: deny\n")
    extra = policy_io.parse_safety_policy("version: v2\ndefault_tenant: t\ntenants:\n  t:\n    allow_topics: [job.b]\n    max_concurrent_jobs: 3\n    mcp:\n      deny_tools: [dd]\n  u:\n    deny_topics: [job.x]\nrules:\n  - id: e\n    decision: allow\n")
    m = policy_io.merge_policies(base, extra)
    assert m["version"] == "v1" and m["default_tenant"] == "t"                       # base wins when set, else the fragment's
    assert [r["id"] for r in m["rules"]] == ["b", "e"]                               # fragment rules after base rules
    assert m["tenants"]["t"]["allow_topics"] == ["job.a", "job.b"] and m["tenants"]["t"]["max_concurrent_jobs"] == 3   # the smaller positive limit
    assert m["tenants"]["t"]["mcp"]["deny_tools"] == ["rm", "dd"] and m["tenants"]["u"]["deny_topics"] == ["job.x"]
    assert policy_io.merge_policies(None, extra)["rules"][0]["id"] == "e" and policy_io.merge_policies(base, None)["version"] == "v1"
    base["rules"][0]["id"] = "changed"                                               # the merge result is a deep copy
    assert m["rules"][0]["id"] == "b"


def test_json_merge_patch():   # RFC 7386 (pack overlays on pools / timeouts)
    t = {"a": 1, "b": {"c": 2, "d": 3}, "e": [1, 2]}
    assert policy_io.json_merge_patch(t, {"b": {"c": None, "x": 9}, "e": [3], "f": "g"}) == {"a": 1, "b": {"d": 3, "x": 9}, "e": [3], "f": "g"}
    assert policy_io.json_merge_patch(t, 7) == 7 and t["b"]["c"] == 2
