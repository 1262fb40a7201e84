"""Small end-to-end run of every kernel for compute-sanitizer (tools/sanitize.sh): policy_kernel, route_kernel (label
picks, multi-chunk and unsorted pools), worker_chunk / worker_merge, the device encoder, heartbeat ingest, scheduler ticks,
and the wide-mask path.  Checks the results against the oracle, so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kats  # noqa: E402
import oracle_lib  # noqa: E402
from cordum_b200 import engine, synth, wire  # noqa: E402

FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")


def same(got, want, what):
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), (what, f)
    print("ok:", what, flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    light = len(sys.argv) > 2 and sys.argv[2] == "light"   # racecheck serialises everything: smaller pools, fewer phases
    cfg = synth.make_config("c2", n)                      # 256 rules, 1k workers
    e = engine.Engine(device=0)
    e.load_policy(cfg.policy, "san")
    e.load_routing(cfg.routing)
    e.load_workers(cfg.workers)
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    b = e.batch(n)
    for mode in (wire.MODE_POLICY_AND_ROUTE, wire.MODE_POLICY_ONLY, wire.MODE_ROUTE_ONLY):
        same(b.encode(cfg.jobs).dispatch(mode), o.eval(cfg.jobs, mode), "c2 host-encode mode %d" % mode)
    same(b.encode_device(cfg.jobs).dispatch(), o.eval(cfg.jobs), "c2 device-encode")
    # heartbeat deltas + refresh
    rng = np.random.default_rng(3)
    slots = rng.choice(cfg.workers.n_workers, 200, replace=False).astype(np.uint32)
    loads = np.zeros(200, wire.LOAD_DTYPE)
    loads["active_jobs"] = rng.integers(0, 4, 200)
    loads["max_parallel_jobs"] = 4
    loads["cpu_load"] = rng.integers(0, 95, 200)
    e.update_workers(slots, loads)
    o.update_workers(slots, loads)
    same(b.encode(cfg.jobs).dispatch(), o.eval(cfg.jobs), "after heartbeat deltas")
    # one big pool: multi-chunk merge (1300) and beyond the sort buffer (9000)
    for size in ((1300,) if light else (1300, 9000)):
        routing = {"topics": {"job.w.go": ["big"]}, "pools": {"big": {}}}
        workers = [kats.hb("m%05d" % i, "big", i % 3, float(i % 40), float(i % 7), 4, {"zone": "z%d" % (i % 4), "host": "m%05d" % (i % 97)}) for i in range(size)]
        e.load_routing(routing)
        e.load_workers(workers)
        o2 = oracle_lib.Oracle(cfg.policy, routing, workers)
        jobs = [{"topic": "job.w.go", "labels": {"zone": "z%d" % (i % 4)}} for i in range(64)] + \
               [{"topic": "job.w.go", "labels": {"host": "m%05d" % (i % 97), "zone": "z%d" % (i % 4)}} for i in range(64)] + [{"topic": "job.w.go"}] * 8
        same(b.encode(jobs).dispatch(wire.MODE_ROUTE_ONLY), o2.eval(jobs, wire.MODE_ROUTE_ONLY), "pool of %d workers" % size)
        o2.close()
    # wide masks
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_wide_masks as tw

    policy, routing, workers, jobs = tw.build(70 if light else 130)
    e.load_policy(policy, "wide")
    e.load_routing(routing)
    e.load_workers(workers)
    o3 = oracle_lib.Oracle(policy, routing, workers)
    bw = e.batch(len(jobs))
    same(bw.encode(jobs).dispatch(), o3.eval(jobs), "wide masks")
    o3.close()
    # scheduler ticks on the c2 tables
    e.load_policy(cfg.policy, "san2")
    e.load_routing(cfg.routing)
    e.load_workers(cfg.workers)
    o4 = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    b1, b2 = e.batch(n), e.batch(n)
    b1.encode(cfg.jobs).dispatch()
    b2.encode_device(cfg.jobs).dispatch()
    W = cfg.workers.n_workers
    keep = []
    for t in range(4):
        loads = cfg.workers.loads()
        loads["active_jobs"] = rng.integers(0, 9, W)
        loads["cpu_load"] = (rng.random(W) * 100).astype(np.float32)
        keep.append(loads)
        (b1 if t % 2 == 0 else b2).tick(loads.ctypes.data, 0, W)
    b2.wait()                                  # routed by the flush, on the loads of the last tick
    o4.update_workers(np.arange(W, dtype=np.uint32), keep[3])
    same(b2.fetch(), o4.eval(cfg.jobs), "ticks")
    o4.close()
    o.close()
    print("sanitize target done", flush=True)


if __name__ == "__main__":
    main()
