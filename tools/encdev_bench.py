"""Device-encoder benchmark on the B200: c3, de-interned envelopes in pinned memory.
Prints: encode_device (H2D + 3 kernels) ms per batch, and the pipelined end-to-end rate with two batches in flight."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cordum_b200 import engine, synth, wire  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    cfg = synth.make_config("c3", n)
    eng = engine.Engine(0)
    eng.load_policy(cfg.policy, "kb")
    eng.load_routing(cfg.routing)
    eng.load_workers(cfg.workers)
    plain = cfg.jobs.deinterned()
    pins = [eng.pinned_envelopes(plain) for _ in range(2)]
    bs = [eng.batch(n) for _ in range(2)]
    ref = bs[0].encode(cfg.jobs).dispatch().copy()          # host encode: registers the vocabulary
    bytes_in = len(plain.arena) + sum(v.nbytes for k, v in plain.cols.items())
    print("envelope bytes per batch: %.1f MB (%.0f B/job)" % (bytes_in / 1e6, bytes_in / n))
    for rep in range(3):
        t0 = time.perf_counter()
        bs[0].encode_device(pins[0])
        bs[0].wait()
        t1 = time.perf_counter()
        print("encode_device (pinned) + wait: %.2f ms" % ((t1 - t0) * 1e3))
    got = bs[0].dispatch()
    assert np.array_equal(got["decision"], ref["decision"]) and np.array_equal(got["rule_idx"], ref["rule_idx"]) and np.array_equal(got["worker_slot"], ref["worker_slot"])
    print("host fallbacks:", eng.host_fallbacks())
    for label, src in (("pinned", pins), ("pageable", [plain, plain])):
        steps = 10
        for k in range(2):
            bs[k].encode_device(src[k]); bs[k].dispatch_async()
        for b in bs:
            b.wait()
        t0 = time.perf_counter()
        for k in range(steps):
            b = bs[k % 2]
            b.encode_device(src[k % 2])
            b.dispatch_async()
        for b in bs:
            b.wait()
        dt = time.perf_counter() - t0
        print("e2e %s: %.2f ms per 1M-job batch -> %.1f M decisions/s" % (label, dt / steps * 1e3, n * steps / dt / 1e6))


if __name__ == "__main__":
    main()
