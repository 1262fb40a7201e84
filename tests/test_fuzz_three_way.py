"""Adversarial randomized agreement of three independent implementations on small configurations built
from awkward strings (padding, case, glob classes / escapes / malformed patterns, empty values, duplicates,
MCP aliases, effective configs, legacy tenant rules, ties, overloaded workers):

    C++ oracle (oracle/oracle.cpp)  ==  Python oracle (oracle/py_oracle.py)  ==  the product's host tables re-walked
                                                                                 the way the kernels walk them

The reference's own tests pin none of these interactions (last row); agreement of two restatements
written from the cited lines, plus the product's table compiler + encoder, is what stands in for them.  CPU only."""
import json
import os
import random
import sys

import numpy as np
import pytest

import kats
import oracle_lib
import table_walk
from cordum_b200 import wire

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import py_oracle  # noqa: E402

STATUS_CODE = {v: k for k, v in kats.ROUTE_NAMES.items()}
FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")

TENANTS = ["default", "t1", "T1", " t1 ", "acme", "Acme ", ""]
PACKS = ["sre", "db", "Web", "x"]
VERBS = ["collect", "delete", "a", "a.b", "*", ""]
CAPS = ["cap.read", "CAP.READ", " cap.write", "cap.write", ""]
RISKS = ["write", "Write ", "prod", "secrets", "SECRETS", "pii", ""]
REQS = ["git", "GIT", " net ", "gpu", "docker", ""]
LKEYS = ["env", "zone", "tier", "gpu"]
LVALS = ["prod", "dev", "a", "", "true"]
ACTORS = ["alice", "Alice", "svc-1", ""]
POOLS = ["p0", "p1", "p2", "gpu-pool"]
PATTERNS = ["job.*", "job.sre.*", "job.[sd]*.*", "job.[^s]*.collect", "job.db.delete", "job.?b.*", "job.\\*.x", "job.[abc",
            " job.sre.collect ", "", "*", "job.sre.c*t", "JOB.sre.*", "job.web.*", "job.*.a.b"]
MCP_VALS = ["srv", "SRV", " srv ", "other", "bad", "read", "WRITE", ""]
# non-ASCII: Unicode White_Space that TrimSpace strips (NBSP, EM SPACE, IDEOGRAPHIC SPACE, NEL, LINE SEPARATOR), characters it
# does not (ZERO WIDTH SPACE, BOM), letters that only fold outside ASCII (EqualFold is restated as ASCII-only), multi-byte
# runes under '?' and character classes of path.Match
TENANTS += ["\u00a0t1\u00a0", "\u2003acme", "t1\u3000", "\u00c9cole", "\u00e9cole", "t1\u200b", "\u0085t1\u2028"]
CAPS += ["\u00a0cap.read", "CAP.R\u00c9AD", "cap.r\u00e9ad"]
RISKS += ["\u2002write", "write\ufeff"]
REQS += ["\u00a0git", "g\u0131t"]
MCP_VALS += ["\u2003srv\u2003", "s\u0280v"]
PATTERNS += ["job.\u00e9*", "\u00a0job.sre.*\u00a0", "job.[\u00e0-\u00ff]*.*", "job.?\u00e9.*"]
PACKS += ["\u00e9t\u00e9"]
DECISIONS = ["allow", "deny", "require_approval", "throttle", "allow_with_constraints", "permit", "block", "REQUIRE-HUMAN", ""]


def pick(rng, xs, lo=0, hi=2):
    return [rng.choice(xs) for _ in range(rng.randint(lo, hi))]


def mcp_lists(rng):
    m = {}
    for k in ("servers", "tools", "resources", "actions"):
        if rng.random() < 0.25:
            m["allow_" + k] = pick(rng, MCP_VALS, 1, 2)
        if rng.random() < 0.2:
            m["deny_" + k] = pick(rng, MCP_VALS, 1, 2)
    return m


def make_policy(rng, max_rules=12):
    if max_rules <= 12 and rng.random() < 0.15:   # legacy form: no rules, per-tenant allow / deny topic lists
        tenants = {}
        for t in rng.sample(["default", "t1", "acme"], rng.randint(1, 3)):
            tenants[t] = {"allow_topics": pick(rng, PATTERNS, 0, 2), "deny_topics": pick(rng, PATTERNS, 0, 2), "mcp": mcp_lists(rng)}
        return {"default_tenant": rng.choice(["default", " t1", ""]), "tenants": tenants}
    rules = []
    for i in range(rng.randint(max(1, max_rules // 3), max_rules)):
        m = {}
        if rng.random() < 0.5:
            m["tenants"] = pick(rng, TENANTS, 1, 2)
        if rng.random() < 0.7:
            m["topics"] = pick(rng, PATTERNS, 1, 3)
        if rng.random() < 0.3:
            m["capabilities"] = pick(rng, CAPS, 1, 2)
        if rng.random() < 0.3:
            m["risk_tags"] = pick(rng, RISKS, 1, 2)
        if rng.random() < 0.25:
            m["requires"] = pick(rng, REQS, 1, 2)
        if rng.random() < 0.2:
            m["pack_ids"] = pick(rng, PACKS, 1, 2)
        if rng.random() < 0.2:
            m["actor_ids"] = pick(rng, ACTORS, 1, 2)
        if rng.random() < 0.2:
            m["actor_types"] = pick(rng, ["human", "Service", "", "robot"], 1, 2)
        if rng.random() < 0.25:
            m["labels"] = {rng.choice(LKEYS): rng.choice(LVALS) for _ in range(rng.randint(1, 2))}
        if rng.random() < 0.15:
            m["secrets_present"] = rng.random() < 0.5
        if rng.random() < 0.2:
            m["mcp"] = mcp_lists(rng)
        r = {"id": "r%d" % i, "decision": rng.choice(DECISIONS), "reason": "because %d" % i, "match": m}
        if rng.random() < 0.2:
            r["constraints"] = {"budgets": {"max_runtime_ms": 1000 + i}}
        rules.append(r)
    pol = {"default_tenant": rng.choice(["default", "t1", ""]), "rules": rules}
    if rng.random() < 0.4:
        pol["tenants"] = {t: {"mcp": mcp_lists(rng)} for t in rng.sample(["default", "t1", "acme", "T1"], rng.randint(1, 2))}
    return pol


def make_routing(rng):
    topics = {}
    for p in PACKS + [""]:
        for v in VERBS[:4]:
            if rng.random() < 0.6:
                topics["job.%s.%s" % (p.lower(), v)] = [rng.choice(POOLS) for _ in range(rng.randint(1, 3))]
    topics[" job.sre.collect "] = ["p2"]          # the routing map is keyed by the RAW topic
    pools = {p: ({"requires": pick(rng, REQS, 0, 2)} if rng.random() < 0.5 else {}) for p in POOLS if rng.random() < 0.9}
    return {"topics": topics, "pools": pools}


def make_workers(rng):
    ws = []
    for i in range(rng.randint(0, 24)):
        labels = {k: rng.choice(LVALS) for k in rng.sample(LKEYS, rng.randint(0, 3))}
        coarse = rng.random() < 0.6   # coarse loads: ties and exact overload thresholds
        cpu = float(rng.choice([0, 25, 50, 89.5, 90, 95])) if coarse else rng.random() * 100
        gpu = float(rng.choice([0, 0, 50, 90])) if coarse else rng.random() * 100
        ws.append(kats.hb("w%02d" % rng.randint(0, 30), rng.choice(POOLS + ["orphan"]), rng.randint(0, 9), cpu, gpu,
                          rng.choice([0, 4, 10]), labels))
    return ws


def make_job(rng, workers):
    pack, verb = rng.choice(PACKS), rng.choice(VERBS)
    topic = rng.choice(["job.%s.%s" % (pack.lower(), verb), "job.%s.%s" % (pack, verb), " job.sre.collect ", "", "jobs.x", "job.", "job.sre.collect"])
    j = {"topic": topic, "tenant": rng.choice(TENANTS), "principal_id": rng.choice(ACTORS)}
    if rng.random() < 0.8:
        j["meta"] = {"tenant_id": rng.choice(TENANTS), "actor_id": rng.choice(ACTORS), "actor_type": rng.choice([0, 1, 2, 3]),
                     "capability": rng.choice(CAPS), "pack_id": rng.choice(PACKS + [""]),
                     "risk_tags": pick(rng, RISKS, 0, 3), "requires": pick(rng, REQS, 0, 2)}
    labels = {}
    for k in rng.sample(LKEYS, rng.randint(0, 3)):
        labels[k] = rng.choice(LVALS)
    if rng.random() < 0.3:
        labels[rng.choice(["mcp.server", "mcp_server", "mcpServer"])] = rng.choice(MCP_VALS)
    if rng.random() < 0.2:
        labels[rng.choice(["mcp.tool", "mcpTool"])] = rng.choice(MCP_VALS)
    if rng.random() < 0.15:
        labels[rng.choice(["mcp.action", "mcp_action"])] = rng.choice(MCP_VALS)
    if rng.random() < 0.1:
        labels["mcp.resource"] = rng.choice(MCP_VALS)
    if rng.random() < 0.2:
        labels["secrets_present"] = rng.choice(["true", "TRUE", "1", "yes", "YES", "no", " true ", ""])
    if rng.random() < 0.15:
        labels["preferred_pool"] = rng.choice(POOLS + ["nope", ""])
    if rng.random() < 0.15 and workers:
        labels["preferred_worker_id"] = rng.choice([w["worker_id"] for w in workers] + ["ghost"])
    if rng.random() < 0.3:
        labels.update({"workflow_id": "wf", "run_id": "r", "cordum.trace": "x"})
    if labels or rng.random() < 0.5:
        j["labels"] = labels
    if rng.random() < 0.2:
        eff = {}
        if rng.random() < 0.6:
            eff["denied_topics"] = pick(rng, PATTERNS, 1, 2)
        if rng.random() < 0.5:
            eff["allowed_topics"] = pick(rng, PATTERNS, 0, 2)
        if rng.random() < 0.4:
            eff["mcp"] = mcp_lists(rng)
        # ParseEffectiveSafety (effective.go:12-39): encoding/json field matching is case-insensitive, wrong types fail
        shape = rng.random()
        if shape < 0.7:
            j["effective_config"] = json.dumps({"safety": eff})
        elif shape < 0.8:
            j["effective_config"] = json.dumps({"Safety": {k.upper(): v for k, v in eff.items()}})
        elif shape < 0.9:
            j["effective_config"] = json.dumps({"safety": {"denied_topics": "job.*"}})   # not an array
        else:
            j["effective_config"] = rng.choice(["{", "null", "[]", '{"safety": null}', '{"other": 1}', '{"safety": {"denied_topics": [1]}}'])
    if rng.random() < 0.1:
        j["approved"] = True
    return j


def py_records(policy, routing, workers, jobs):
    live = list({w["worker_id"]: w for w in workers}.values())   # registry semantics: the latest heartbeat wins
    out = []
    for job in jobs:
        r = py_oracle.process_job(policy, routing, live, job)
        rt = r["route"]
        out.append((wire.DEC_NAMES.index(r["decision"]), wire.DEC_NAMES.index(r["sched_decision"]), r["rule_idx"],
                    bool(r["approval_required"]), bool(r["has_snapshot"]), bool(r["has_constraints"]),
                    STATUS_CODE[rt["status"]] if rt else 0, bool(rt["tie"]) if rt else False))
    return out


def rec_tuples(rec):
    return [(int(r["decision"]), int(r["sched_decision"]), int(r["rule_idx"]), bool(r["flags"] & wire.F_APPROVAL_REQUIRED),
             bool(r["flags"] & wire.F_HAS_SNAPSHOT), bool(r["flags"] & wire.F_CONSTRAINTS), int(r["route_status"]),
             bool(r["flags"] & wire.F_TIE)) for r in rec]


@pytest.mark.parametrize("seed,max_rules", [(s, 12) for s in range(60)] + [(100 + s, 400) for s in range(12)])
def test_three_implementations_agree(seed, max_rules):
    """max_rules = 400: several 128-bit words of rule positions, so clustering by topic prefix, duplicated positions of
    multi-pattern rules, per-topic word lists and the ascending order inside a word all take part."""
    rng = random.Random(7000 + seed)
    policy, routing, workers = make_policy(rng, max_rules), make_routing(rng), make_workers(rng)
    jobs = [make_job(rng, workers) for _ in range(40 if max_rules <= 12 else 120)]
    env = wire.EnvelopeBatch.from_jobs(jobs)
    o = oracle_lib.Oracle(policy, routing, workers)
    h = table_walk.HostHarness(policy, routing, workers)
    for mode in (wire.MODE_POLICY_AND_ROUTE, wire.MODE_POLICY_ONLY, wire.MODE_ROUTE_ONLY):
        want_m, got = o.eval(env, mode), h.evaluate(env, mode)
        for f in FIELDS:   # product host tables vs C++ oracle: every field of the record
            bad = np.nonzero(got[f] != want_m[f])[0]
            assert len(bad) == 0, "seed %d mode %d field %s job %s: tables %s oracle %s\n%s" % (
                seed, mode, f, bad[:4], got[f][bad[:4]], want_m[f][bad[:4]], json.dumps(jobs[int(bad[0])]))
    want = o.eval(env, wire.MODE_POLICY_AND_ROUTE)
    o.close()
    h.close()
    a, b = rec_tuples(want), py_records(policy, routing, workers, jobs)
    bad = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert not bad, "seed %d C++ vs Python oracle: %s\n%s" % (seed, bad[:3], json.dumps(jobs[bad[0][0]]))


@pytest.mark.parametrize("seed", range(30))
def test_reloads_and_heartbeats_keep_agreeing(seed):
    """The same engine-side objects across a life cycle: heartbeat deltas (loads by slot), a worker-registry reload, a
    routing reload and a policy reload, each followed by a re-evaluation of freshly encoded jobs - dictionaries are
    rebuilt, ids are reassigned and the tables must still agree with an oracle that went through the same calls."""
    rng = random.Random(9000 + seed)
    policy, routing, workers = make_policy(rng), make_routing(rng), make_workers(rng)
    o = oracle_lib.Oracle(policy, routing, workers)
    h = table_walk.HostHarness(policy, routing, workers)

    def check(what):
        jobs = [make_job(rng, workers) for _ in range(30)]
        env = wire.EnvelopeBatch.from_jobs(jobs)
        want, got = o.eval(env, wire.MODE_POLICY_AND_ROUTE), h.evaluate(env, wire.MODE_POLICY_AND_ROUTE)
        for f in FIELDS:
            bad = np.nonzero(got[f] != want[f])[0]
            assert len(bad) == 0, "seed %d after %s: field %s job %s: tables %s oracle %s\n%s" % (
                seed, what, f, bad[:4], got[f][bad[:4]], want[f][bad[:4]], json.dumps(jobs[int(bad[0])]))

    check("load")
    for step in range(6):
        kind = rng.choice(["heartbeat", "heartbeat", "workers", "routing", "policy"])
        if kind == "heartbeat" and workers:
            n = len(workers)
            slots = np.array(sorted(rng.sample(range(n), rng.randint(1, n))), dtype=np.uint32)
            loads = np.zeros(len(slots), dtype=wire.LOAD_DTYPE)
            loads["active_jobs"] = [rng.randint(0, 9) for _ in slots]
            loads["max_parallel_jobs"] = [rng.choice([0, 4, 10]) for _ in slots]
            loads["cpu_load"] = np.array([rng.choice([0, 50, 89.99, 90, 100 * rng.random()]) for _ in slots], dtype=np.float32)
            loads["gpu_utilization"] = np.array([rng.choice([0, 0, 90, 100 * rng.random()]) for _ in slots], dtype=np.float32)
            o.update_workers(slots, loads)
            h.update_workers(slots, loads)
        elif kind == "workers":
            workers = make_workers(rng)
            o.load_workers(workers)
            h.load_workers(workers)
        elif kind == "routing":
            routing = make_routing(rng)
            o.load_routing(routing)
            h.load_routing(routing)
        elif kind == "policy":
            policy = make_policy(rng)
            o.load_policy(policy)
            h.load_policy(policy)
        check("%s (step %d)" % (kind, step))
    o.close()
    h.close()


@pytest.mark.parametrize("seed", range(16))
def test_three_implementations_agree_with_wide_universes(seed, monkeypatch):
    """The same agreement with vocabularies large enough that the risk-tag, requires-token, label-pair and placement-label
    dictionaries outgrow the mask fields of the job record (tables.h WideLayout: side rows of extra mask words), mixed
    with the awkward strings above, and across a reload back to a narrow policy."""
    me = sys.modules[__name__]
    monkeypatch.setattr(me, "RISKS", RISKS + ["tag%02d" % i for i in range(90)] + ["TAG%02d " % i for i in range(0, 90, 7)])
    monkeypatch.setattr(me, "REQS", REQS + ["req%02d" % i for i in range(100)] + [" REQ%02d" % i for i in range(0, 100, 9)])
    monkeypatch.setattr(me, "LKEYS", LKEYS + ["k%02d" % i for i in range(30)])
    monkeypatch.setattr(me, "LVALS", LVALS + ["v%d" % i for i in range(8)])
    rng = random.Random(12000 + seed)
    policy, routing, workers = make_policy(rng, 300), make_routing(rng), make_workers(rng) + make_workers(rng) + make_workers(rng)
    for r in policy.get("rules", []):   # more rules with mask predicates than the base generator makes
        if rng.random() < 0.5:
            r["match"]["risk_tags"] = pick(rng, me.RISKS, 2, 5)
        if rng.random() < 0.4:
            r["match"]["requires"] = pick(rng, me.REQS, 1, 3)
        if rng.random() < 0.4:
            r["match"]["labels"] = {rng.choice(me.LKEYS): rng.choice(me.LVALS) for _ in range(rng.randint(1, 2))}
    for p in routing["pools"].values():
        if rng.random() < 0.7:
            p["requires"] = pick(rng, me.REQS, 1, 6)
    h = table_walk.HostHarness(policy, routing, workers)
    T = h.tables()
    assert T["xw_risk"] + T["xw_req"] + T["xw_lab"] + T["xw_place"] >= 2, (T["xw_risk"], T["xw_req"], T["xw_lab"], T["xw_place"])
    o = oracle_lib.Oracle(policy, routing, workers)
    jobs = []
    for _ in range(150):
        j = make_job(rng, workers)
        if rng.random() < 0.5:
            j.setdefault("meta", {})
            j["meta"] = dict(j["meta"] or {})
            j["meta"]["risk_tags"] = pick(rng, me.RISKS, 0, 5)
            j["meta"]["requires"] = pick(rng, me.REQS, 0, 5)
        jobs.append(j)
    env = wire.EnvelopeBatch.from_jobs(jobs)
    for mode in (wire.MODE_POLICY_AND_ROUTE, wire.MODE_POLICY_ONLY, wire.MODE_ROUTE_ONLY):
        want_m, got = o.eval(env, mode), h.evaluate(env, mode)
        for f in FIELDS:
            bad = np.nonzero(got[f] != want_m[f])[0]
            assert len(bad) == 0, "seed %d mode %d field %s job %s: tables %s oracle %s\n%s" % (
                seed, mode, f, bad[:4], got[f][bad[:4]], want_m[f][bad[:4]], json.dumps(jobs[int(bad[0])]))
    a, b = rec_tuples(o.eval(env, wire.MODE_POLICY_AND_ROUTE)), py_records(policy, routing, workers, jobs)
    bad = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert not bad, "seed %d C++ vs Python oracle: %s\n%s" % (seed, bad[:3], json.dumps(jobs[bad[0][0]]))
    # narrow again: the side rows disappear and the same envelopes still agree
    monkeypatch.undo()
    narrow = make_policy(random.Random(seed), 12)
    h.load_policy(narrow)
    o2 = oracle_lib.Oracle(narrow, routing, workers)
    got, want = h.evaluate(env), o2.eval(env)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), (seed, f)
    o.close(); o2.close(); h.close()
