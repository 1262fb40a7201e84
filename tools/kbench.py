"""Kernel micro-benchmark for iteration on the B200: c3, device-resident, prints dispatch_kernel ms.
usage: python tools/kbench.py [reps] ; honours CORDUM_* tuning env vars read by the library."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cordum_b200 import engine, synth, wire  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    name = os.environ.get("KBENCH_CONFIG", "c3")
    cfg = synth.make_config(name)
    eng = engine.Engine(0)
    eng.load_policy(cfg.policy, "kb")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    nb = 4
    bs = [eng.batch(cfg.jobs.n_jobs) for _ in range(nb)]
    t0 = time.time()
    for b in bs:
        b.encode(cfg.jobs)
        b.dispatch()
    enc = (time.time() - t0) / nb
    for mode, nm in ((wire.MODE_POLICY_AND_ROUTE, "policy+route"), (wire.MODE_POLICY_ONLY, "policy"), (wire.MODE_ROUTE_ONLY, "route")):
        ks, ps, rs = [], [], []
        for r in range(reps + 2):
            b = bs[r % nb]
            b.dispatch_resident(mode)
            if r >= 2:
                ks.append(b.timing()[1])
                p_, r_ = b.kernel_times()
                ps.append(p_)
                rs.append(r_)
        print("%s %-13s kernels_ms med %.4f (min %.4f)  policy %.4f  route %.4f" % (
            name, nm, float(np.median(ks)), min(ks), float(np.median(ps)), float(np.median(rs))), flush=True)
    print("encode+dispatch s/batch %.4f" % enc)


if __name__ == "__main__":
    main()
