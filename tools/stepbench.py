"""Host-side cost of one bench step (enqueue only) vs GPU time, small shard: python tools/stepbench.py [jobs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cordum_b200 import engine, synth, wire  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
cfg = synth.make_config("c3")
jobs = cfg.jobs.slice(0, n)
W = cfg.workers.n_workers
eng = engine.Engine(0)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
bs = [eng.batch(n) for _ in range(8)]
for b in bs:
    b.encode(jobs)
    b.dispatch()
loads = torch.from_numpy(cfg.workers.loads().view(np.uint8).reshape(-1, 16).copy()).pin_memory()
dev = torch.empty((W, 16), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
for variant in ("full", "no-ingest"):
    torch.cuda.synchronize()
    K = 200
    t0 = time.perf_counter()
    th = 0.0
    for k in range(K):
        a = time.perf_counter()
        if variant == "full":
            dev.copy_(loads, non_blocking=True)
            eng.set_loads_device(dev.data_ptr(), W, stream.cuda_stream)
        bs[k % 8].dispatch_resident_async()
        th += time.perf_counter() - a
    for b in bs:
        b.wait()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s: %d jobs/step  step %.1f us  host enqueue %.1f us  kernels %s" % (variant, n, dt / K * 1e6, th / K * 1e6, bs[0].kernel_times()))
