"""Policies and registries whose dictionaries outgrow the mask fields of the job record: more than 64 distinct risk tags /
requires tokens / label pairs referenced by rules and pools, more than 128 placement-label bits on routable workers
(tables.h WideLayout).  The further bits travel in a side row of 64-bit words per job; results stay bit-exact.
Reference semantics: safety_policy.go:308-345 (containsAny / containsAll / labelsMatch), strategy_least_loaded.go:161-175
(matchesLabels), :241-265 (poolSatisfies).

CPU: the host tables + encoder, walked in Python (table_walk), against the oracle.  GPU: the kernels against the oracle."""
import random

import numpy as np
import pytest

import kats
import oracle_lib
import table_walk
from cordum_b200 import wire

FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")


def assert_same(got, want, what=""):
    assert len(got) == len(want)
    for f in FIELDS:
        bad = np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, "%s field %s differs at %d jobs, first %s: got %s want %s" % (
            what, f, len(bad), bad[:8], got[f][bad[:8]], want[f][bad[:8]])


def build(n: int, seed: int = 1):
    """n distinct risk tags, requires tokens, label pairs and placement labels, all of them decisive for some job."""
    rng = random.Random(seed * 1000 + n)
    tags = ["tag%03d" % i for i in range(n)]
    reqs = ["Req%03d" % i for i in range(n)]
    pairs = [("k%02d" % (i % 37), "v%03d" % i) for i in range(n)]
    rules = []
    for i in range(n):   # one rule per tag / token set / pair, in an order that interleaves them
        rules.append({"id": "risk-%d" % i, "decision": "deny", "reason": "r", "match": {"topics": ["job.r.*"], "risk_tags": [tags[i], tags[(i * 7 + 3) % n]]}})
        rules.append({"id": "req-%d" % i, "decision": "require_approval", "reason": "q",
                      "match": {"topics": ["job.q.*"], "requires": [reqs[i], reqs[(i + n // 2) % n].lower()]}})
        rules.append({"id": "lab-%d" % i, "decision": "throttle", "reason": "l",
                      "match": {"topics": ["job.l.*"], "labels": dict([pairs[i], pairs[(i * 5 + 1) % n]] if pairs[i][0] != pairs[(i * 5 + 1) % n][0] else [pairs[i]])}})
    rules.append({"id": "empty-pair", "decision": "deny", "reason": "e", "match": {"topics": ["job.e.*"], "labels": {"absent-key": ""}}})
    policy = {"default_tenant": "default", "rules": rules}
    # pools declare overlapping slices of the requires universe; every worker carries a label nobody else has + shared ones
    n_pools = 6
    pools = {"pool%d" % p: {"requires": [reqs[i].upper() for i in range(n) if i % n_pools == p or i % 3 == 0]} for p in range(n_pools)}
    routing = {"topics": {"job.w.go": list(pools), "job.q.x": list(pools), "job.r.x": ["pool0"], "job.l.x": ["pool1"], "job.e.x": ["pool2"]}, "pools": pools}
    workers = []
    for i in range(n):
        labels = {"host": "h%03d" % i, "zone": "z%d" % (i % 5), "serial": "s%03d" % (i * 7 % n)}
        if i % 4 == 0:
            labels["rack-%d" % i] = ""          # a key of its own with an empty value
        workers.append(kats.hb("w%03d" % i, "pool%d" % (i % n_pools), i % 3, float(i % 50), float((i * 3) % 40), 4, labels))
    jobs = []
    for i in range(4 * n):
        k = (i * 7 + i // 8) % n
        kind = i % 8
        if kind == 0:
            jobs.append({"topic": "job.r.x", "meta": {"risk_tags": [tags[k]]}})
        elif kind == 1:
            jobs.append({"topic": "job.r.x", "meta": {"risk_tags": rng.sample(tags, 5) + ["unknown"]}})
        elif kind == 2:
            jobs.append({"topic": "job.q.x", "meta": {"requires": [reqs[k].upper(), reqs[(k + n // 2) % n]]}})
        elif kind == 3:
            jobs.append({"topic": "job.q.x", "meta": {"requires": rng.sample(reqs, min(n, 9))}})
        elif kind == 4:
            a, b = pairs[k], pairs[(k * 5 + 1) % n]
            jobs.append({"topic": "job.l.x", "labels": dict([a, b])})
        elif kind == 5:
            jobs.append({"topic": "job.l.x", "labels": dict(rng.sample(pairs, 6))})
        elif kind == 6:
            lab = {"host": "h%03d" % k} if k % 3 else {"serial": "s%03d" % (k * 7 % n)}
            if k % 2:
                lab["zone"] = "z%d" % (k % 5 if k % 3 else (k + 1) % 5)
            if k % 4 == 0 and k % 8 == 0:
                lab["rack-%d" % k] = ""
            jobs.append({"topic": "job.w.go", "labels": lab})
        else:
            req = [reqs[j] for j in range(n) if j % 6 == k % 6 and rng.random() < 0.5][:12] or [reqs[k]]
            jobs.append({"topic": "job.w.go", "meta": {"requires": req}, "labels": {"zone": "z%d" % (k % 5), "preferred_worker_id": "w%03d" % k}})
    jobs.append({"topic": "job.e.x"})
    jobs.append({"topic": "job.e.x", "labels": {"absent-key": "x"}})
    jobs.append({"topic": "job.e.x", "labels": {"other": "x"}})
    return policy, routing, workers, jobs


@pytest.mark.parametrize("n", [64, 65, 128, 200])
def test_host_tables_with_wide_masks_match_the_oracle(n):
    policy, routing, workers, jobs = build(n)
    h = table_walk.HostHarness(policy, routing, workers)
    T = h.tables()
    if n == 64:
        assert (T["xw_risk"], T["xw_req"]) == (0, 0) and T["xw_lab"] == 1        # 64 pairs + the ("absent-key","") pair
    else:
        assert T["xw_risk"] == (n + 63) // 64 - 1 and T["xw_req"] >= (n + 63) // 64 - 1 and T["xw_lab"] >= (n + 63) // 64 - 1
    assert T["xw_place"] == max(0, (T["place_bits"] + 63) // 64 - 2) and (T["place_bits"] > 128 or n == 64)
    o = oracle_lib.Oracle(policy, routing, workers)
    for mode in (wire.MODE_POLICY_AND_ROUTE, wire.MODE_POLICY_ONLY, wire.MODE_ROUTE_ONLY):
        want = o.eval(jobs, mode)
        assert_same(h.evaluate(jobs, mode), want, "n=%d mode=%d" % (n, mode))
    # the wide bits are decisive: some jobs match rules / workers only through bits beyond the record's own
    want = o.eval(jobs)
    hit = set(want["rule_idx"].tolist())
    assert len(hit) > n // 2 and (want["route_status"] == wire.ROUTE_OK).sum() > n // 4
    if n > 64:   # rules 3*i .. 3*i+2 belong to tag / token / pair i: some are matched through bits beyond 63
        assert any(r >= 3 * 64 for r in hit)
    # a reload back to a narrow policy drops the side rows again
    h.load_policy({"rules": [{"id": "x", "decision": "deny", "match": {"risk_tags": ["tag001"]}}]})
    h.load_routing({"topics": {"job.w.go": ["pool0"]}, "pools": {"pool0": {}}})
    h.load_workers(workers[:6])
    T = h.tables()
    assert (T["xw_risk"], T["xw_req"], T["xw_lab"], T["xw_place"]) == (0, 0, 0, 0)
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("device_encode", [False, True], ids=["host-encode", "device-encode"])
@pytest.mark.parametrize("n", [65, 128, 200])
def test_kernels_with_wide_masks_match_the_oracle(n, device_encode):
    from cordum_b200 import engine

    policy, routing, workers, jobs = build(n)
    e = engine.Engine(device=0)
    e.load_policy(policy, "wide")
    e.load_routing(routing)
    e.load_workers(workers)
    o = oracle_lib.Oracle(policy, routing, workers)
    b = e.batch(len(jobs))
    for mode in (wire.MODE_POLICY_AND_ROUTE, wire.MODE_POLICY_ONLY, wire.MODE_ROUTE_ONLY):
        enc = b.encode_device(jobs) if device_encode else b.encode(jobs)   # the device encoder hands wide tables to the host's
        assert_same(enc.dispatch(mode), o.eval(jobs, mode), "n=%d mode=%d" % (n, mode))
    if device_encode:
        assert e.host_fallbacks() >= 3
    # back to a narrow policy on the same engine and batch
    narrow = {"rules": [{"id": "x", "decision": "deny", "reason": "x", "match": {"risk_tags": ["tag001"]}}]}
    e.load_policy(narrow, "narrow")
    o2 = oracle_lib.Oracle(narrow, routing, workers)
    assert_same(b.encode(jobs).dispatch(), o2.eval(jobs), "narrow again")
    # and a many-worker pool: the label bitmaps of a multi-chunk pool carry bits beyond 128 too (worker_merge_kernel)
    big = [kats.hb("m%04d" % i, "pool0", i % 2, float(i % 30), 0.0, 4, {"host": "m%04d" % i, "zone": "z%d" % (i % 3)}) for i in range(1300)]
    e.load_workers(big)
    o3 = oracle_lib.Oracle(narrow, routing, big)
    pick = [{"topic": "job.w.go", "labels": {"host": "m%04d" % i}} for i in range(0, 1300, 37)] + \
           [{"topic": "job.w.go", "labels": {"zone": "z1", "host": "m%04d" % i}} for i in range(0, 1300, 41)]
    assert_same(b.encode(pick).dispatch(), o3.eval(pick), "multi-chunk pool")
    b.free()
    e.close()
    o.close(); o2.close(); o3.close()
