"""Tick time on one GPU for the per-rank shard sizes of 1/2/4/8-GPU runs (no peer exchange: the slice is the whole table).
usage: python tools/tickbench.py [ticks]   ; honours CORDUM_TICK_SHARE"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cordum_b200 import engine, synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = synth.make_config("c3")
W = cfg.workers.n_workers
eng = engine.Engine(0)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
rng = np.random.default_rng(1)
sets = []
for s in range(4):
    d = cfg.workers.loads()
    d["active_jobs"] = rng.integers(0, 9, W)
    d["cpu_load"] = (rng.random(W) * 100).astype(np.float32)
    sets.append(torch.from_numpy(d.view(np.uint8).reshape(-1, 16).copy()).pin_memory())
for n in (1_000_000, 125_000):
    jobs = cfg.jobs.slice(0, n)
    n_rot = min(6, max(2, int(np.ceil(2.0 * (126 << 20) / (n * 112))) + 1))
    bs = [eng.batch(n) for _ in range(n_rot)]
    for b in bs:
        b.encode(jobs).dispatch()
    ts = torch.cuda.ExternalStream(eng.tick_stream)
    for k in range(6 * n_rot):      # every (batch pair, phase) graph is captured outside the timed loop
        bs[k % n_rot].tick(sets[k % 4].data_ptr(), 0, W)
    eng.tick_flush()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(ts)
    for k in range(K):
        bs[k % n_rot].tick(sets[k % 4].data_ptr(), 0, W)
    eng.tick_flush()
    b_.record(ts)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / K
    print("share %s  jobs/tick %7d  tick %.1f us  -> %.2f G decisions/s per GPU" % (os.environ.get("CORDUM_TICK_SHARE", "2"), n, ms * 1e3, n / ms / 1e6), flush=True)
    for b in bs:
        b.free()
