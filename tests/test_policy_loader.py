"""The safety kernel's policy loader and reload loop (kernel.go:485-881) over the engine.  The CPU tests follow the
reference's own: kernel_test.go:147-223,284-320, policy_source_test.go:23-101, helpers_test.go:82-92.  The GPU tests swap
policies under the engine while requests are in flight."""
import base64
import hashlib
import http.server
import sys
import threading

import pytest

import oracle_lib
from cordum_b200 import policy_io, policy_loader as pl, wire

sys.path.insert(0, oracle_lib.ORACLE_DIR)
import py_oracle  # noqa: E402

TEST_POLICY = b"version: v1\ndefault_tenant: default\ntenants:\n  default:\n    allow_topics:\n      - job.*\n"   # policy_source_test.go:15-21

BUNDLES = {   # kernel_test.go:173-196
    "alpha": "\ndefault_tenant: default\ntenants:\n  default:\n    allow_topics:\n      - job.*\n",
    "beta": {"content": "\nrules:\n  - id: require-prod\n    match:\n      topics:\n        - job.prod.*\n    decision: require_approval\n    reason: prod writes\n"},
    "disabled": {"content": "tenants:\n  default:\n    deny_topics:\n      - job.disabled\n", "enabled": False},
}


def test_policy_source_from_env(monkeypatch):
    monkeypatch.setenv("SAFETY_POLICY_URL", "http://example")
    assert pl.policy_source_from_env("/tmp/policy.yaml") == "http://example"
    monkeypatch.delenv("SAFETY_POLICY_URL")
    assert pl.policy_source_from_env("/tmp/policy.yaml") == "/tmp/policy.yaml"


def test_load_policy_bundle_from_file(tmp_path):
    path = tmp_path / "policy.yaml"
    path.write_bytes(TEST_POLICY)
    policy, snapshot = pl.load_policy_bundle(str(path))
    assert policy["version"] == "v1" and snapshot == "v1:" + hashlib.sha256(TEST_POLICY).hexdigest()
    unversioned = TEST_POLICY.replace(b"version: v1\n", b"")
    path.write_bytes(unversioned)
    assert pl.load_policy_bundle(str(path))[1] == hashlib.sha256(unversioned).hexdigest()
    assert pl.load_policy_bundle("") == (None, "")


def test_read_policy_source_http():
    class H(http.server.BaseHTTPRequestHandler):
        def do_GET(self):
            self.send_response(200 if self.path == "/ok" else 503)
            self.end_headers()
            self.wfile.write(TEST_POLICY)

        def log_message(self, *a):
            pass
    srv = http.server.HTTPServer(("127.0.0.1", 0), H)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        assert pl.read_policy_source("http://127.0.0.1:%d/ok" % srv.server_port) == TEST_POLICY
        with pytest.raises(Exception):
            pl.read_policy_source("http://127.0.0.1:%d/bad" % srv.server_port)
    finally:
        srv.shutdown()


def test_verify_policy_signature(monkeypatch, tmp_path):
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey

    priv = Ed25519PrivateKey.generate()
    pub = priv.public_key().public_bytes(serialization.Encoding.Raw, serialization.PublicFormat.Raw)
    sig = priv.sign(TEST_POLICY)
    pl.verify_policy_signature(TEST_POLICY, "policy.yaml")                       # no key configured: nothing to verify
    monkeypatch.setenv("SAFETY_POLICY_PUBLIC_KEY", base64.b64encode(pub).decode())
    monkeypatch.setenv("SAFETY_POLICY_SIGNATURE", base64.b64encode(sig).decode())
    pl.verify_policy_signature(TEST_POLICY, "policy.yaml")
    with pytest.raises(ValueError, match="verification failed"):
        pl.verify_policy_signature(TEST_POLICY + b" ", "policy.yaml")
    monkeypatch.delenv("SAFETY_POLICY_SIGNATURE")
    with pytest.raises(ValueError, match="no signature provided"):
        pl.verify_policy_signature(TEST_POLICY, "https://example/policy.yaml")
    with pytest.raises(ValueError, match="not found"):
        pl.verify_policy_signature(TEST_POLICY, str(tmp_path / "p.yaml"))
    (tmp_path / "p.yaml").write_bytes(TEST_POLICY)
    (tmp_path / "p.yaml.sig").write_bytes(sig)                                   # <source>.sig next to the file
    assert pl.load_policy_bundle(str(tmp_path / "p.yaml"))[1].startswith("v1:")
    monkeypatch.setenv("SAFETY_POLICY_SIGNATURE_PATH", str(tmp_path / "p.yaml.sig"))
    pl.verify_policy_signature(TEST_POLICY, "elsewhere.yaml")
    monkeypatch.setenv("SAFETY_POLICY_PUBLIC_KEY", "@@@")
    with pytest.raises(ValueError, match="invalid SAFETY_POLICY_PUBLIC_KEY"):
        pl.verify_policy_signature(TEST_POLICY, "policy.yaml")


def test_decode_key():
    assert pl.decode_key(base64.b64encode(b"hello").decode()) == b"hello"
    assert pl.decode_key(b"hello".hex()) == b"hello"
    with pytest.raises(ValueError):
        pl.decode_key("@@@")
    with pytest.raises(ValueError):
        pl.decode_key("")


def test_parse_bool_and_combine_snapshots():
    assert pl.parse_bool("yes") and not pl.parse_bool("no") and pl.parse_bool(" ON ") and not pl.parse_bool("")
    assert pl.combine_snapshots("a", "") == "a" and pl.combine_snapshots("a", "b") == "a|b" and pl.combine_snapshots("", "b") == "b"


def test_extract_policy_fragment_honors_enabled():
    assert pl.extract_policy_fragment({"content": "foo", "enabled": True}) == ("foo", True)
    assert pl.extract_policy_fragment({"content": "bar", "enabled": False}) == ("", False)
    assert pl.extract_policy_fragment({"policy": "p", "enabled": "Yes"}) == ("p", True)
    assert pl.extract_policy_fragment({"data": "d", "enabled": 1.0}) == ("d", True)
    assert pl.extract_policy_fragment({"data": "d", "enabled": 0}) == ("", False)
    assert pl.extract_policy_fragment({"other": "x"}) == ("", False) and pl.extract_policy_fragment(7) == ("", False)


def test_policy_loader_loads_fragments():
    loader = pl.PolicyLoader(bundles=lambda: BUNDLES)
    policy, snapshot = loader.load_fragments()
    h = hashlib.sha256()
    for key in ("alpha", "beta"):   # sorted keys; the disabled bundle is not hashed
        content = BUNDLES[key] if isinstance(BUNDLES[key], str) else BUNDLES[key]["content"]
        h.update(key.encode() + b"\x00" + content.encode())
    assert snapshot == "cfg:" + h.hexdigest()
    pd = py_oracle.policy_evaluate(policy, {"tenant": "default", "topic": "job.prod.test", "labels": {}, "meta": py_oracle.policy_meta({}),
                                            "mcp": py_oracle.extract_mcp({}), "secrets_present": False})
    assert pd["decision"] == "require_approval"
    assert "job.disabled" not in str(policy)
    assert pl.PolicyLoader(bundles=lambda: None).load_fragments() == (None, "")
    assert pl.PolicyLoader(bundles=lambda: {"x": {"content": "  ", "enabled": True}}).load_fragments() == (None, "")
    with pytest.raises(ValueError, match='parse policy fragment "bad"'):
        pl.PolicyLoader(bundles=lambda: {"bad": "rules: [ {decision: nonsense} ]"}).load_fragments()


def test_policy_loader_from_source_and_defaults(tmp_path):
    path = tmp_path / "policy.yaml"
    path.write_bytes(b"default_tenant: default\ntenants:\n  default:\n    allow_topics:\n      - job.*\n")
    policy, snapshot = pl.PolicyLoader(source=str(path)).load()
    assert policy is not None and snapshot != ""
    assert not pl.PolicyLoader().should_watch() and pl.PolicyLoader(source="/tmp/policy.yaml").should_watch()
    # base + fragments: base rules first, snapshots joined with '|'
    policy, snapshot = pl.PolicyLoader(source=str(path), bundles=lambda: BUNDLES).load()
    assert snapshot.count("|") == 1 and snapshot.split("|")[1].startswith("cfg:")
    assert [r["id"] for r in policy["rules"]] == ["require-prod"] and policy["default_tenant"] == "default"


def test_reload_interval(monkeypatch):
    for raw, want in (("", 30.0), ("5s", 5.0), ("250ms", 0.25), ("2m", 120.0), ("-1s", 30.0), ("junk", 30.0)):
        monkeypatch.setenv("SAFETY_POLICY_RELOAD_INTERVAL", raw)
        assert pl.reload_interval_from_env() == want


# ---------------------------------------------------------------------------------------------------- GPU
def _policy(n_rules: int, tag: str) -> dict:
    rules = [{"id": "%s-%d" % (tag, i), "decision": "deny", "reason": tag, "match": {"topics": ["job.t%d.*" % i]}} for i in range(n_rules)]
    return {"default_tenant": "default", "rules": rules}


@pytest.mark.gpu
def test_watcher_swaps_the_policy_when_the_snapshot_changes(tmp_path):
    from cordum_b200 import reference_api as api

    path = tmp_path / "policy.yaml"
    path.write_bytes(TEST_POLICY)
    srv = api.SafetyKernelServer()
    logs = []
    w = pl.PolicyWatcher(srv, pl.PolicyLoader(source=str(path), bundles=lambda: BUNDLES), interval_s=3600, log=logs.append)
    assert w.poll_once() and not w.poll_once()                       # second poll: same snapshot, nothing to do
    r = srv.check({"job_id": "j", "topic": "job.prod.x", "tenant": "default"})
    assert r["decision"] == "REQUIRE_HUMAN" and r["policy_snapshot"].startswith("v1:") and "|cfg:" in r["policy_snapshot"]
    path.write_bytes(TEST_POLICY + b"rules:\n  - id: stop\n    decision: deny\n    reason: stop\n    match:\n      topics: [job.prod.*]\n")
    assert w.poll_once()
    r2 = srv.check({"job_id": "j", "topic": "job.prod.x", "tenant": "default"})
    assert r2["decision"] == "DENY" and r2["rule_id"] == "stop" and r2["policy_snapshot"] != r["policy_snapshot"]
    path.write_bytes(b"rules: [ {decision: nonsense} ]")              # a broken reload keeps the policy in force
    assert not w.poll_once() and "reload failed" in logs[-1]
    assert srv.check({"topic": "job.prod.x", "tenant": "default"})["rule_id"] == "stop"
    assert srv.list_snapshots()[0] == r2["policy_snapshot"] and len(srv.list_snapshots()) == 2


@pytest.mark.gpu
def test_requests_racing_a_reload_are_answered_under_one_policy_or_the_other():
    """watchPolicy swaps s.policy under the write lock while evaluate holds the read lock (kernel.go:140-147,510-521): a
    request sees policy A with snapshot A, or policy B with snapshot B.  Here: the front-end re-encodes a batch whose
    tables were swapped between encode and dispatch, and reports the snapshot the batch ran under."""
    from cordum_b200 import engine, frontend

    eng = engine.Engine(device=0)
    pa, pb = _policy(300, "A"), _policy(300, "B")
    pb["rules"][7]["decision"] = "allow"
    eng.load_policy(pa, "snap-A")
    eng.load_routing({"topics": {}, "pools": {}})
    eng.load_workers([])
    fe = frontend.Frontend(eng, max_batch=64, max_wait_us=50, mode=wire.MODE_POLICY_ONLY)
    stop = threading.Event()
    bad, seen = [], {"snap-A": 0, "snap-B": 0}

    def client(k):
        i = 0
        while not stop.is_set():
            t = (k * 131 + i) % 300
            r = fe.submit({"topic": "job.t%d.x" % t, "tenant": "default"})
            snap = r.snapshot.decode()
            want_rule = ("A-%d" if snap == "snap-A" else "B-%d") % t
            want_dec = wire.DEC_ALLOW if (snap == "snap-B" and t == 7) else wire.DEC_DENY
            if r.status != 0 or snap not in seen or r.rule_id.decode() != want_rule or r.rec.decision != want_dec:
                bad.append((r.status, snap, r.rule_id.decode(), int(r.rec.decision), r.reason.decode()))
            else:
                seen[snap] += 1
            i += 1

    threads = [threading.Thread(target=client, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for i in range(40):
        eng.load_policy(pb if i % 2 == 0 else pa, "snap-B" if i % 2 == 0 else "snap-A")
    stop.set()
    for t in threads:
        t.join()
    fe.close()
    eng.close()
    assert not bad, bad[:5]
    assert seen["snap-A"] > 0 and seen["snap-B"] > 0
