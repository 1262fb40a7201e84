"""Policy hot-reload cost on the c3 policy (4096 rules): what a watchPolicy swap (kernel.go:485-521) costs here.
  load     cordum_policy_load: JSON parse + table compile (host), epoch bump
  encode   first encode of a 64k-job batch afterwards (encoder caches are invalidated by the reload)
  first    first dispatch afterwards (table upload happens here: sync_tables) vs the steady-state dispatch
  stall    while one thread reloads the policy every `period` ms, the front-end's request latency (p50 / p99 / max)
usage: python tools/reloadbench.py [out.json]"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cordum_b200 import engine, frontend, hostinfo, policy_io, synth, wire  # noqa: E402


def main():
    cfg = synth.make_config("c3", 65536)
    eng = engine.Engine(0)
    eng.load_policy(cfg.policy, "r0")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    doc = policy_io.to_json(cfg.policy)
    b = eng.batch(cfg.jobs.n_jobs)
    b.encode(cfg.jobs).dispatch()
    steady = []
    for _ in range(20):
        t0 = time.perf_counter()
        b.dispatch()
        steady.append((time.perf_counter() - t0) * 1e3)
    rows = []
    for i in range(12):
        t0 = time.perf_counter()
        eng.load_policy(doc, "r%d" % (i + 1))
        t1 = time.perf_counter()
        b.encode(cfg.jobs)
        t2 = time.perf_counter()
        b.dispatch()
        t3 = time.perf_counter()
        b.dispatch()
        t4 = time.perf_counter()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
    a = np.asarray(rows[2:])
    out = {"host": hostinfo.describe(), "workload": "c3 policy: %d rules, %d bytes of JSON; batch of %d jobs" % (len(cfg.policy["rules"]), len(doc), cfg.jobs.n_jobs),
           "load_ms_p50": float(np.median(a[:, 0])), "first_encode_ms_p50": float(np.median(a[:, 1])),
           "first_dispatch_ms_p50": float(np.median(a[:, 2])), "next_dispatch_ms_p50": float(np.median(a[:, 3])),
           "steady_dispatch_ms_p50": float(np.median(steady))}
    out["reload_to_first_decision_ms"] = out["load_ms_p50"] + out["first_encode_ms_p50"] + out["first_dispatch_ms_p50"]
    print("policy reload, ms (median of 10): load %.2f | first encode %.2f | first dispatch %.2f (then %.2f; steady %.2f) | reload -> first decisions %.2f"
          % (out["load_ms_p50"], out["first_encode_ms_p50"], out["first_dispatch_ms_p50"], out["next_dispatch_ms_p50"], out["steady_dispatch_ms_p50"],
             out["reload_to_first_decision_ms"]), flush=True)

    # request latency through the front-end while the policy is swapped every `period` ms
    jobs = cfg.jobs.to_jobs(0, 4096)
    for period_ms in (0, 1000, 100):
        fe = frontend.Frontend(eng, max_batch=256, max_wait_us=100, mode=wire.MODE_POLICY_AND_ROUTE)
        stop = threading.Event()
        lat = [[] for _ in range(16)]
        fails = [0]

        def client(k):
            i = k
            while not stop.is_set():
                t0 = time.perf_counter()
                r = fe.submit(jobs[i % len(jobs)])
                lat[k].append((time.perf_counter() - t0) * 1e6)
                if r.status != 0:
                    fails[0] += 1
                i += 16

        ts = [threading.Thread(target=client, args=(k,)) for k in range(16)]
        for t in ts:
            t.start()
        t_end = time.perf_counter() + 4.0
        n_reloads = 0
        while time.perf_counter() < t_end:
            if period_ms:
                eng.load_policy(doc, "s%d" % n_reloads)
                n_reloads += 1
                time.sleep(period_ms / 1e3)
            else:
                time.sleep(0.05)
        stop.set()
        for t in ts:
            t.join()
        fe.close()
        xs = np.concatenate([np.asarray(x) for x in lat])
        row = {"reload_period_ms": period_ms, "reloads": n_reloads, "requests": int(xs.size), "failed": fails[0],
               "p50_us": float(np.percentile(xs, 50)), "p99_us": float(np.percentile(xs, 99)), "max_us": float(xs.max())}
        out.setdefault("under_reload", []).append(row)
        print("front-end, 16 python client threads, reload every %4d ms: %d reloads, %d requests (%d failed)  p50 %.0f us  p99 %.0f us  max %.0f us"
              % (period_ms, n_reloads, xs.size, fails[0], row["p50_us"], row["p99_us"], row["max_us"]), flush=True)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)
    b.free()
    eng.close()


if __name__ == "__main__":
    main()
