"""Host encoder scaling (no GPU work): python tools/encbench.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import table_walk  # noqa: E402
from cordum_b200 import synth  # noqa: E402

cfg = synth.make_config("c3")
n = cfg.jobs.n_jobs
print("cpus", os.cpu_count())
for th in [int(x) for x in (sys.argv[1:] or ["1", "8", "16", "32", "64"])]:
    h = table_walk.HostHarness(cfg.policy, cfg.routing, cfg.workers, threads=th)
    L = h.L
    slab = np.zeros(int(L.cordum_test_slab_bytes(n)) + 16, dtype=np.uint8)
    best = 1e9
    for rep in range(5):
        t = time.perf_counter()
        L.cordum_test_host_encode(h.h, C.addressof(cfg.jobs.struct), slab.ctypes.data)
        best = min(best, time.perf_counter() - t)
    print("threads %3d  best %.4f s  %.1f Mjobs/s  %.0f ns/job/thread" % (th, best, n / best / 1e6, best / n * 1e9 * th), flush=True)
    h.close()
