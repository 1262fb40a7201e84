"""Where does the end-to-end step time go?  python tools/e2ebench.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
from cordum_b200 import engine, synth, wire  # noqa: E402

print("affinity", len(os.sched_getaffinity(0)))
cfg = synth.make_config("c3")
eng = engine.Engine(0)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
bs = [eng.batch(cfg.jobs.n_jobs) for _ in range(2)]
for b in bs:
    b.encode(cfg.jobs)
    b.dispatch()
for it in range(6):
    b = bs[it % 2]
    t0 = time.perf_counter()
    b.encode(cfg.jobs)
    t1 = time.perf_counter()
    b.dispatch_async()
    t2 = time.perf_counter()
    print("step %d encode(+wait prev) %.2f ms  enqueue %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
for b in bs:
    b.wait()
t0 = time.perf_counter()
bs[0].encode(cfg.jobs)
print("encode alone %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
bs[0].dispatch()
print("dispatch (H2D+kernels+D2H) blocking %.2f ms, events total %.2f ms" % ((time.perf_counter() - t0) * 1e3, bs[0].timing()[0]))
