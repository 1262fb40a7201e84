"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle.
Bit-exact on every field of the decision record.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import kats
import oracle_lib
from cordum_b200 import synth, wire

pytestmark = pytest.mark.gpu

FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")


def assert_same(got, want, what=""):
    assert len(got) == len(want)
    for f in FIELDS:
        bad = np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, "%s field %s differs at %d jobs, first %s: got %s want %s" % (
            what, f, len(bad), bad[:8], got[f][bad[:8]], want[f][bad[:8]])


@pytest.fixture(scope="module", params=["host-encode", "device-encode"])
def eng(request):
    """Every test of this module runs twice: with the host encoder (cordum_encode) and with the device encoder
    (cordum_encode_device: the envelope bytes go to the GPU as they are and are dictionary-coded there)."""
    from cordum_b200 import engine

    engine.Batch.default_device_encode = request.param == "device-encode"
    e = engine.Engine(device=0)
    yield e
    e.close()
    engine.Batch.default_device_encode = False


def load(e, policy, routing, workers):
    e.load_policy(policy, "test")
    e.load_routing(routing)
    e.load_workers(workers)


@pytest.mark.parametrize("case", kats.CASES, ids=[c["name"] for c in kats.CASES])
def test_reference_kats_through_c_abi(eng, case):
    load(eng, case["policy"], case["routing"], case["workers"])
    b = eng.batch(8)
    rec = b.encode([case["job"]]).dispatch(case["mode"])[0]
    got = {
        "decision": wire.DEC_NAMES[rec["decision"]], "sched_decision": wire.DEC_NAMES[rec["sched_decision"]],
        "reason": b.reason(0), "rule_id": eng.rule_id(int(rec["rule_idx"])), "rule_idx": int(rec["rule_idx"]),
        "approval_required": bool(rec["flags"] & wire.F_APPROVAL_REQUIRED),
        "has_snapshot": bool(rec["flags"] & wire.F_HAS_SNAPSHOT), "has_constraints": bool(rec["flags"] & wire.F_CONSTRAINTS),
        "route": kats.ROUTE_NAMES[int(rec["route_status"])], "subject": b.subject(0),
        "worker_slot": int(rec["worker_slot"]), "tie": bool(rec["flags"] & wire.F_TIE)}
    kats.check(case, got)
    want = oracle_lib.Oracle(case["policy"], case["routing"], case["workers"]).eval([case["job"]], case["mode"])
    assert_same(b.results(), want, case["name"])
    b.free()


@pytest.mark.parametrize("name,n,mode", [("tiny", 2000, wire.MODE_POLICY_AND_ROUTE), ("tiny", 2000, wire.MODE_POLICY_ONLY),
                                         ("tiny", 2000, wire.MODE_ROUTE_ONLY), ("c2", None, wire.MODE_POLICY_AND_ROUTE),
                                         ("c2", None, wire.MODE_POLICY_ONLY), ("c2", None, wire.MODE_ROUTE_ONLY)])
def test_synthetic_configs_bit_exact(eng, name, n, mode):
    cfg = synth.make_config(name, n)
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    b = eng.batch(cfg.jobs.n_jobs)
    got = b.encode(cfg.jobs).dispatch(mode)
    want = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers).eval(cfg.jobs, mode, threads=8)
    assert_same(got, want, "%s mode %d" % (name, mode))
    b.free()


@pytest.mark.parametrize("lbest", ["0", "1"])
def test_label_picks_with_and_without_the_label_best_table(eng, monkeypatch, lbest):
    """Single-label jobs are answered from the per-(pool, label) table the refresh builds, or by the bitmap scan: the engine
    chooses by world size (engine.cu view()); CORDUM_LBEST forces either.  Both are exact."""
    monkeypatch.setenv("CORDUM_LBEST", lbest)
    cfg = synth.make_config("c2", 6000)
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    b = eng.batch(cfg.jobs.n_jobs)
    for mode in (wire.MODE_POLICY_AND_ROUTE, wire.MODE_ROUTE_ONLY):
        assert_same(b.encode(cfg.jobs).dispatch(mode), o.eval(cfg.jobs, mode, threads=8), "lbest=%s mode %d" % (lbest, mode))
    loads = cfg.workers.loads()
    loads["active_jobs"] = (np.arange(len(loads)) * 7) % 9
    slots = np.arange(len(loads), dtype=np.uint32)
    eng.update_workers(slots, loads)
    o.update_workers(slots, loads)
    assert_same(b.dispatch(), o.eval(cfg.jobs, threads=8), "lbest=%s after a heartbeat epoch" % lbest)
    b.free()
    o.close()


def test_ragged_batch_sizes(eng):
    cfg = synth.make_config("tiny", 300)
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    want = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers).eval(cfg.jobs, wire.MODE_POLICY_AND_ROUTE)
    b = eng.batch(300)
    for n in (1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 255, 300):
        got = b.encode(cfg.jobs.slice(0, n)).dispatch()
        assert_same(got, want[:n], "n=%d" % n)
    b.free()


def test_ties_and_heartbeat_deltas(eng):
    spec = synth.Spec(**{**synth.SPECS["tiny"].__dict__, "tie_fraction": 0.5, "seed": 11})
    cfg = synth.make_config(spec)
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    b = eng.batch(cfg.jobs.n_jobs)
    want = o.eval(cfg.jobs)
    assert (want["flags"] & wire.F_TIE).astype(bool).sum() > 20
    assert_same(b.encode(cfg.jobs).dispatch(), want, "ties")
    rng = np.random.default_rng(1)
    for it in range(3):
        slots = rng.permutation(cfg.workers.n_workers)[: cfg.workers.n_workers // 2].astype(np.uint32)
        loads = np.zeros(len(slots), dtype=wire.LOAD_DTYPE)
        loads["active_jobs"] = rng.integers(0, 6, len(slots))
        loads["max_parallel_jobs"] = rng.choice([0, 4, 8], len(slots))
        loads["cpu_load"] = (rng.random(len(slots)) * 100).astype(np.float32)
        loads["gpu_utilization"] = (rng.random(len(slots)) * 100).astype(np.float32)
        eng.update_workers(slots, loads)
        o.update_workers(slots, loads)
        assert_same(b.dispatch(), o.eval(cfg.jobs), "after heartbeat deltas %d" % it)
    b.free()


def test_device_load_table_path(eng):
    """cordum_workers_set_loads_device: the multi-GPU exchange hands the gathered load table over in HBM."""
    import torch

    cfg = synth.make_config("tiny")
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    rng = np.random.default_rng(3)
    loads = cfg.workers.loads()
    loads["active_jobs"] = rng.integers(0, 9, len(loads))
    loads["cpu_load"] = (rng.random(len(loads)) * 100).astype(np.float32)
    o.update_workers(np.arange(len(loads), dtype=np.uint32), loads)
    dev = torch.from_numpy(loads.view(np.uint8).reshape(-1, 16).copy()).cuda()
    torch.cuda.synchronize()
    eng.set_loads_device(dev.data_ptr(), len(loads), torch.cuda.current_stream().cuda_stream)
    b = eng.batch(cfg.jobs.n_jobs)
    assert_same(b.encode(cfg.jobs).dispatch(), o.eval(cfg.jobs), "device loads")
    b.free()


def test_engine_owned_ingest_single_rank(eng):
    """cordum_workers_ingest without an exchange (world = 1): the slice is the whole registry; it must be this rank's
    share exactly, epochs alternate between two gather buffers, and results follow the ingested loads."""
    cfg = synth.make_config("tiny")
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    rng = np.random.default_rng(4)
    n = cfg.workers.n_workers
    b = eng.batch(cfg.jobs.n_jobs)
    b.encode(cfg.jobs)
    keep = []
    for epoch in range(5):
        loads = cfg.workers.loads()
        loads["active_jobs"] = rng.integers(0, 9, n)
        loads["cpu_load"] = (rng.random(n) * 100).astype(np.float32)
        loads["gpu_utilization"] = (rng.random(n) * 100).astype(np.float32)
        keep.append(loads)                       # the host buffer must stay alive until the copy has run
        eng.ingest(loads.ctypes.data, 0, n)
        o.update_workers(np.arange(n, dtype=np.uint32), loads)
        assert_same(b.dispatch(), o.eval(cfg.jobs), "ingest epoch %d" % epoch)
    with pytest.raises(Exception):
        eng.ingest(keep[0].ctypes.data, 0, n - 1)
    with pytest.raises(Exception):
        eng.ingest(keep[0].ctypes.data, 1, n)
    b.free()


@pytest.mark.parametrize("n_big,coarse", [(512, False), (513, True), (1500, True), (5000, False), (8192, True), (8800, False)])
def test_pool_sizes_across_refresh_paths(eng, n_big, coarse):
    """Worker-table refresh paths by pool size: one chunk (<= 512 workers, finished by worker_chunk_kernel), several
    chunks merged by worker_merge_kernel (<= 8192), and the unsorted pool whose jobs scan (route_kernel S').
    `coarse` loads make many equal scores, so the merge must keep worker_id order inside a score and report ties."""
    rng = np.random.default_rng(n_big)
    n = n_big + 200

    def mk(i):
        cpu, gpu = (float(rng.integers(0, 3)) * 25.0, 0.0) if coarse else (float(rng.random() * 100), float(rng.random() * 100))
        return kats.hb("w%05d" % i, "big" if i < n_big else "small", int(rng.integers(0, 3 if coarse else 6)), cpu, gpu,
                       int(rng.choice([0, 4, 8])),
                       {"zone": "z%d" % rng.integers(0, 4), **({"gpu": "true"} if rng.random() < 0.3 else {})})
    workers = [mk(i) for i in range(n)]
    rng.shuffle(workers)   # slot order is not worker_id order
    routing = {"topics": {"job.a": ["big", "small"], "job.b": ["small"], "job.c": ["big"]}, "pools": {"big": {}, "small": {}}}
    jobs = []
    for i in range(400):
        labels = {}
        if i % 3 != 0:
            labels["zone"] = "z%d" % (i % 5)       # z4 matches nobody
        if i % 4 == 0:
            labels["gpu"] = "true"
        jobs.append({"topic": ("job.a", "job.b", "job.c")[i % 3] if i % 5 else "job.b", "labels": labels})
    load(eng, None, routing, workers)
    b = eng.batch(len(jobs))
    got = b.encode(jobs).dispatch(wire.MODE_ROUTE_ONLY).copy()
    o = oracle_lib.Oracle(None, routing, workers)
    assert_same(got, o.eval(jobs, wire.MODE_ROUTE_ONLY), "pool of %d" % n_big)
    assert set(got["route_status"].tolist()) >= {wire.ROUTE_OK, wire.ROUTE_NO_WORKERS}
    if coarse:
        assert (got["flags"] & wire.F_TIE).any()
    # a heartbeat epoch that overloads most of the big pool, then one that overloads all of it
    for frac in (0.9, 1.0):
        slots = np.arange(n, dtype=np.uint32)
        loads = wire.WorkerTable.from_workers(workers).loads()
        hot = rng.random(n) < frac
        loads["cpu_load"] = np.where(hot, 95.0, loads["cpu_load"]).astype(np.float32)
        eng.update_workers(slots, loads)
        o.update_workers(slots, loads)
        assert_same(b.dispatch(wire.MODE_ROUTE_ONLY), o.eval(jobs, wire.MODE_ROUTE_ONLY), "pool of %d, %.0f%% overloaded" % (n_big, frac * 100))
    b.free()


def test_many_risk_tags(eng):
    """Jobs with up to seven referenced risk tags (policy_kernel reads three risk rows branch-free, the rest loop)."""
    rng = np.random.default_rng(12)
    tags = ["t%d" % i for i in range(10)]
    rules = [{"id": "r%d" % i, "decision": ["deny", "require_approval", "throttle"][i % 3], "reason": "x",
              "match": {"topics": ["job.a.*"], "risk_tags": [tags[i % 10], tags[(i * 3 + 1) % 10]]}} for i in range(40)]
    policy = {"default_tenant": "default", "rules": rules}
    jobs = []
    for n in range(0, 8):
        for _ in range(40):
            jobs.append({"topic": "job.a.x", "meta": {"risk_tags": [str(x) for x in rng.choice(tags, size=n, replace=False)]}})
    load(eng, policy, {"topics": {}, "pools": {}}, [])
    b = eng.batch(len(jobs))
    got = b.encode(jobs).dispatch(wire.MODE_POLICY_ONLY).copy()
    assert_same(got, oracle_lib.Oracle(policy, {"topics": {}, "pools": {}}, []).eval(jobs, wire.MODE_POLICY_ONLY), "many risk tags")
    assert len(set(got["rule_idx"].tolist())) > 5
    b.free()


def test_many_placement_labels(eng):
    """More than four placement labels on a job (route_kernel keeps four bitmap rows in registers, the rest loop)."""
    rng = np.random.default_rng(11)
    keys = ["zone", "tier", "region", "env", "arch", "gpu", "disk"]
    workers = [kats.hb("w%03d" % i, "p", int(rng.integers(0, 4)), float(rng.random() * 80), 0.0, 0,
                       {k: "v%d" % rng.integers(0, 3) for k in keys}) for i in range(300)]
    routing = {"topics": {"job.a": ["p"]}, "pools": {"p": {}}}
    jobs = [{"topic": "job.a", "labels": {k: "v%d" % rng.integers(0, 3) for k in keys[:n]}} for n in range(0, 8) for _ in range(20)]
    load(eng, None, routing, workers)
    b = eng.batch(len(jobs))
    got = b.encode(jobs).dispatch(wire.MODE_ROUTE_ONLY).copy()
    assert_same(got, oracle_lib.Oracle(None, routing, workers).eval(jobs, wire.MODE_ROUTE_ONLY), "many labels")
    assert set(got["route_status"].tolist()) >= {wire.ROUTE_OK, wire.ROUTE_NO_WORKERS}
    b.free()


def test_c5_demo_guardrails_replay_100k(eng):
    c5 = kats.golden("c5_demo_guardrails.json")
    jobs, workers, kind = synth.make_c5(100_000)
    load(eng, c5["policy"], c5["routing"], workers)
    o = oracle_lib.Oracle(c5["policy"], c5["routing"], workers)
    b = eng.batch(jobs.n_jobs)
    got = b.encode(jobs).dispatch().copy()
    assert_same(got, o.eval(jobs, threads=8), "c5")
    assert (got["sched_decision"][kind == 0] == wire.DEC_REQUIRE_HUMAN).all()
    assert (got["flags"][kind == 0] & wire.F_APPROVAL_REQUIRED).all()
    assert (got["decision"][kind == 2] == wire.DEC_DENY).all()
    idx = int(np.nonzero(kind == 0)[0][0])
    assert b.reason(idx) == "Write operations require approval." and eng.rule_id(int(got["rule_idx"][idx])) == "demo-guardrails-approval"
    idx = int(np.nonzero(kind == 2)[0][0])
    assert eng.rule_remediations(int(got["rule_idx"][idx]))[0]["id"] == "use-safe"
    replay = jobs.with_approved(kind == 0)
    got2 = b.encode(replay).dispatch()
    assert_same(got2, o.eval(replay, threads=8), "c5 replay")
    assert (got2["route_status"][kind == 0] == wire.ROUTE_OK).all()
    b.free()


def test_dynamic_dictionaries_and_policy_reload(eng):
    policy = {"default_tenant": "default",
              "rules": [{"id": "r", "decision": "deny", "reason": "x", "match": {"topics": ["job.a.*"]}}]}
    routing = {"topics": {"job.b.one": ["p"]}, "pools": {"p": {}}}
    workers = [kats.hb("w", "p")]
    load(eng, policy, routing, workers)
    o = oracle_lib.Oracle(policy, routing, workers)
    jobs = [{"topic": "job.a.new1"}, {"topic": "job.b.one"}, {"topic": " job.a.padded "},
            {"topic": "job.b.one", "effective_config": b'{"safety":{"denied_topics":["job.b.*"]}}'},
            {"topic": "job.b.one", "effective_config": b'{"safety":{"allowed_topics":["job.c.*"]}}'},
            {"topic": "job.c.x", "effective_config": b'{"safety":{"allowed_topics":["job.c.*"]}}'},
            {"topic": "job.b.one", "labels": {"mcp.server": "evil"},
             "effective_config": b'{"data":{"safety":{"mcp":{"deny_servers":["EVIL"]}}}}'},
            {"topic": "job.b.one", "effective_config": b'not json'}]
    b = eng.batch(16)
    for _ in range(2):
        assert_same(b.encode(jobs).dispatch(), o.eval(jobs), "dynamic")
    assert b.reason(3) == "topic 'job.b.one' denied by effective config"
    assert b.reason(6) == 'mcp server "evil" denied'
    # reload: a batch encoded before the reload must be refused, not silently evaluated with stale ids
    from cordum_b200.engine import CordumError

    eng.load_policy({"rules": [{"id": "a", "decision": "allow", "match": {"topics": ["job.a.*"]}}]}, "snap-2")
    with pytest.raises(CordumError):
        b.dispatch()
    assert b.encode(jobs).dispatch()["decision"][0] == wire.DEC_ALLOW
    assert eng.snapshots()[0] == "snap-2"
    b.free()


@pytest.fixture(scope="module")
def c3(eng):
    cfg = synth.make_config("c3")
    load(eng, cfg.policy, cfg.routing, cfg.workers)
    b = eng.batch(cfg.jobs.n_jobs)
    got = b.encode(cfg.jobs).dispatch().copy()
    yield cfg, b, got
    b.free()


def usable_cores() -> int:
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def test_c3_full_batch_bit_exact(eng, c3):
    """1M x 4096 x 65536 (BASELINE config 3, the headline config): EVERY one of the 1,000,000 records the GPU wrote
    is compared with the oracle ("identical to the oracle on every job").  The oracle is
    O(J*(R+W)); on the GPU box's host cores the full batch takes 6-40 s."""
    cfg, b, got = c3
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    want = o.eval(cfg.jobs, wire.MODE_POLICY_AND_ROUTE, threads=usable_cores())
    assert len(want) == cfg.jobs.n_jobs == 1_000_000
    assert_same(got, want, "c3 full batch")
    assert len(set(got["decision"].tolist())) == 5 and len(set(got["route_status"].tolist())) >= 7


def test_c3_full_size_properties(eng, c3):
    """Size-independent properties at the full 1M-job size."""
    cfg, b, got = c3
    # determinism + device-resident path == copy path
    b.dispatch_resident()
    assert_same(b.fetch(), got, "resident rerun")
    # splitting the batch must not change any record (jobs are independent)
    b2 = eng.batch(250_000)
    for k in range(4):
        part = b2.encode(cfg.jobs.slice(k * 250_000, 250_000)).dispatch()
        assert_same(part, got[k * 250_000:(k + 1) * 250_000], "quarter %d" % k)
    b2.free()
    # structural invariants of the record
    routed = np.isin(got["route_status"], (wire.ROUTE_OK, wire.ROUTE_OK_PREFERRED))
    assert ((got["worker_slot"] >= 0) == routed).all()
    allowed = np.isin(got["sched_decision"], (wire.DEC_ALLOW, wire.DEC_ALLOW_WITH_CONSTRAINTS))
    assert (got["route_status"][~allowed] == wire.ROUTE_NOT_ATTEMPTED).all()
    assert (got["route_status"][allowed] != wire.ROUTE_NOT_ATTEMPTED).all()
    assert (got["rule_idx"] < cfg.spec.n_rules).all()
