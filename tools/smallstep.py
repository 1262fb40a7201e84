"""The fixed part of a sharded step, on ONE GPU: the bench step (cordum_workers_ingest of the whole 65,536-worker table +
cordum_dispatch_resident_async) at the per-rank batch sizes of 1/2/4/8-GPU runs.  What does not shrink with the batch is
what bounds strong scaling.  usage: python tools/smallstep.py [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cordum_b200 import engine, synth, wire  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cfg = synth.make_config("c3")
W = cfg.workers.n_workers
eng = engine.Engine(0)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
rng = np.random.default_rng(1)
deltas = []
for i in range(4):
    l = cfg.workers.loads()
    l["active_jobs"] = rng.integers(0, 5, W)
    l["cpu_load"] = (rng.random(W) * 100).astype(np.float32)
    deltas.append(torch.from_numpy(l.view(np.uint8).reshape(-1, 16).copy()).pin_memory())
print("jobs/step | us/step (device) | host enqueue us/step | ingest only us | dispatch only us")
for n in (1_000_000, 500_000, 250_000, 125_000, 32_768):
    jobs = cfg.jobs.slice(0, n)
    bs = [eng.batch(n) for _ in range(4)]
    for b in bs:
        b.encode(jobs).dispatch()
    res = []
    for variant in ("full", "ingest", "dispatch"):
        for b in bs:
            b.wait()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = 0.0
        for k in range(K):
            a = time.perf_counter()
            if variant != "dispatch":
                eng.ingest(deltas[k % 4].data_ptr(), 0, W)
            if variant != "ingest":
                bs[k % 4].dispatch_resident_async(wire.MODE_POLICY_AND_ROUTE | wire.FLAG_NO_TIMING)
            th += time.perf_counter() - a
        for b in bs:
            b.wait()
        torch.cuda.synchronize()
        res.append(((time.perf_counter() - t0) / K * 1e6, th / K * 1e6))
    print("%9d | %8.1f | %8.1f | %8.1f | %8.1f" % (n, res[0][0], res[0][1], res[1][0], res[2][0]), flush=True)
    for b in bs:
        b.free()
