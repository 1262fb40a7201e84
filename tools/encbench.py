"""Host encoder benchmark (no GPU): cordum_encode's work through the host-only hook, interned vs de-interned arena.
usage: python tools/encbench.py [n_jobs] [threads]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import table_walk  # noqa: E402
from cordum_b200 import hostinfo, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print("host", hostinfo.describe())
    cfg = synth.make_config("c3", n)
    h = table_walk.HostHarness(cfg.policy, cfg.routing, cfg.workers, threads=threads)
    L = h.L
    slab = np.zeros(int(L.cordum_test_slab_bytes(n)) + 64, dtype=np.uint8)
    slab = slab[(-slab.ctypes.data) % 16:]
    for name, env in (("interned", cfg.jobs), ("deinterned", cfg.jobs.deinterned())):
        ts = []
        for _ in range(6):
            t0 = time.perf_counter()
            rc = L.cordum_test_host_encode(h.h, C.addressof(env.struct), slab.ctypes.data, None)   # c3 has no wide masks
            ts.append(time.perf_counter() - t0)
            assert rc == 0
        print("%-11s arena %6.1f MB  encode ms: %s  -> best %.1f M jobs/s" % (
            name, len(env.arena) / 1e6, " ".join("%.1f" % (t * 1e3) for t in ts), n / min(ts) / 1e6))


if __name__ == "__main__":
    main()
