"""Per-request latency through the C ABI (the reference publishes p99 4.2 ms for its safety check, BENCHMARKS.md:41-46).
  (1) one batch of n requests: encode + dispatch round trip (host encoder and device encoder), n = 1 ... 1M
  (2) the micro-batching front-end: T client threads, each submitting one request per blocking call
usage: python tools/latency.py [out.json]"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cordum_b200 import engine, frontend, hostinfo, synth, wire  # noqa: E402


def pct(xs, q):
    return float(np.percentile(np.asarray(xs), q))


def main():
    cfg = synth.make_config("c3")
    eng = engine.Engine(0)
    eng.load_policy(cfg.policy, "lat")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    full = eng.batch(cfg.jobs.n_jobs)
    full.encode(cfg.jobs).dispatch()          # vocabulary + tables resident
    out = {"host": hostinfo.describe(), "workload": "c3 policy+route: 4096 rules, 65536 workers", "batch": [], "frontend": []}
    print("batch round trip (encode + H2D + kernels + D2H), ms:  n | host-encoder p50 p99 | device-encoder p50 p99 | us per request (best p50)")
    for n in (1, 64, 1024, 16384, 131072, 1_000_000):
        jobs = cfg.jobs.slice(0, n).deinterned()
        pin = eng.pinned_envelopes(jobs)
        b = eng.batch(n)
        reps = 200 if n <= 16384 else (30 if n <= 131072 else 8)
        th, td = [], []
        for r in range(reps + 3):
            t0 = time.perf_counter()
            b.encode(jobs).dispatch()
            t1 = time.perf_counter()
            b.encode_device(pin).dispatch()
            t2 = time.perf_counter()
            if r >= 3:
                th.append((t1 - t0) * 1e3)
                td.append((t2 - t1) * 1e3)
        row = {"n": n, "host_p50_ms": pct(th, 50), "host_p99_ms": pct(th, 99), "device_p50_ms": pct(td, 50), "device_p99_ms": pct(td, 99)}
        row["us_per_request"] = 1e3 * min(row["host_p50_ms"], row["device_p50_ms"]) / n
        out["batch"].append(row)
        print("%8d | %8.3f %8.3f | %8.3f %8.3f | %.3f" % (n, row["host_p50_ms"], row["host_p99_ms"], row["device_p50_ms"], row["device_p99_ms"], row["us_per_request"]), flush=True)
        b.free()
        pin.free()
    print("front-end (one blocking call per request):  threads | max_batch max_wait_us | requests/s | p50 us  p99 us | mean batch")
    jobs = cfg.jobs.to_jobs(0, 20000)
    packed = [frontend.pack_request(j) for j in jobs]
    import ctypes as C
    reqs = (frontend.Request * len(packed))(*[p[0] for p in packed])     # native client threads (no interpreter lock in the loop)
    for threads, max_batch, wait_us in ((1, 64, 0), (8, 64, 50), (64, 512, 100), (512, 2048, 200), (4096, 8192, 300)):
        fe = frontend.Frontend(eng, max_batch=max_batch, max_wait_us=wait_us)
        cap = 4_000_000
        lat = np.zeros(cap, dtype=np.float32)
        t_start = time.perf_counter()
        done = fe.L.cordum_frontend_loadgen(fe.h, C.addressof(reqs), len(packed), threads, 2.0, lat.ctypes.data, cap)
        dt = time.perf_counter() - t_start
        allv = lat[lat > 0]
        st = fe.stats()
        row = {"threads": threads, "max_batch": max_batch, "max_wait_us": wait_us, "requests_per_s": done / dt,
               "p50_us": pct(allv, 50), "p99_us": pct(allv, 99), "mean_batch": st["requests"] / max(1, st["batches"])}
        out["frontend"].append(row)
        print("%7d | %5d %5d | %10.0f | %8.1f %8.1f | %.1f" % (threads, max_batch, wait_us, row["requests_per_s"], row["p50_us"], row["p99_us"], row["mean_batch"]), flush=True)
        fe.close()
    out["reference_published"] = {"safety_check_p99_ms": 4.2, "source": "BENCHMARKS.md:41-46 of the reference (other hardware, one request per gRPC call)"}
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()
