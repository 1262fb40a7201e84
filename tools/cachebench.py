"""Decision cache of the policy front-end (SAFETY_DECISION_CACHE_TTL, kernel.go:149-162): request latency and rate with
every request a miss (cache off) and every request a hit, native client threads.  usage: python tools/cachebench.py [out.json]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cordum_b200 import engine, frontend, hostinfo, synth, wire  # noqa: E402


def main():
    cfg = synth.make_config("c3")
    eng = engine.Engine(0)
    eng.load_policy(cfg.policy, "cache")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    jobs = cfg.jobs.to_jobs(0, 2000)
    packed = [frontend.pack_request(j) for j in jobs]
    reqs = (frontend.Request * len(packed))(*[p[0] for p in packed])
    out = {"host": hostinfo.describe(), "workload": "c3 policy-only, 2000 distinct requests round-robin", "rows": []}
    print("threads | cache | requests/s | p50 us | p99 us | hits")
    for threads in (1, 16):
        for ttl in (0, 60_000_000):
            fe = frontend.Frontend(eng, max_batch=256, max_wait_us=50, mode=wire.MODE_POLICY_ONLY, cache_ttl_us=ttl)
            cap = 8_000_000
            lat = np.zeros(cap, dtype=np.float32)
            fe.L.cordum_frontend_loadgen(fe.h, C.addressof(reqs), len(packed), threads, 0.3, lat.ctypes.data, cap)   # fill
            lat[:] = 0
            t0 = time.perf_counter()
            done = fe.L.cordum_frontend_loadgen(fe.h, C.addressof(reqs), len(packed), threads, 1.5, lat.ctypes.data, cap)
            dt = time.perf_counter() - t0
            v = lat[lat > 0]
            st = fe.cache_stats()
            row = {"threads": threads, "cache_ttl_us": ttl, "requests_per_s": done / dt, "p50_us": float(np.percentile(v, 50)),
                   "p99_us": float(np.percentile(v, 99)), "hits": st["hits"], "misses": st["misses"]}
            out["rows"].append(row)
            print("%7d | %5s | %10.0f | %7.2f | %7.2f | %d" % (threads, "on" if ttl else "off", row["requests_per_s"], row["p50_us"], row["p99_us"], st["hits"]), flush=True)
            fe.close()
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()
