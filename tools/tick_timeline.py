"""GPU-activity timeline of a few scheduler ticks via torch.profiler (CUPTI).  python tools/tick_timeline.py [jobs] [ticks]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from cordum_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = synth.make_config("c3")
W = cfg.workers.n_workers
eng = engine.Engine(0)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
jobs = cfg.jobs.slice(0, n)
bs = [eng.batch(n) for _ in range(2)]
for b in bs:
    b.encode(jobs).dispatch()
loads = torch.from_numpy(cfg.workers.loads().view(np.uint8).reshape(-1, 16).copy()).pin_memory()
for k in range(12):
    bs[k % 2].tick(loads.data_ptr(), 0, W)
eng.tick_flush()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for k in range(ticks):
        bs[k % 2].tick(loads.data_ptr(), 0, W)
    eng.tick_flush()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "tick_timeline.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
api = [e for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver")]
t0 = min(e["ts"] for e in gpu)
print("GPU activities (us from first): start dur stream name")
for e in sorted(gpu, key=lambda e: e["ts"]):
    print("%9.1f %8.1f  s%-4s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:50]))
print("API:")
for e in sorted(api, key=lambda e: e["ts"])[:60]:
    print("%9.1f %8.1f  %s" % (e["ts"] - t0, e["dur"], e["name"][:40]))
json.dump({"traceEvents": gpu + api}, open(path, "w"))
