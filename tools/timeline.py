"""Kernel/runtime-API timeline of a few bench steps via torch.profiler (CUPTI) — the nsys substitute.
   python tools/timeline.py [jobs] [steps]            (or under torchrun for the sharded step)
Writes gpurun_out/timeline_rank0.json (chrome trace) and prints a per-step table of GPU activities."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from cordum_b200 import engine, shard, synth  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = synth.make_config("c3")
J, W = cfg.jobs.n_jobs, cfg.workers.n_workers
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
jobs = cfg.jobs.slice(0, n)
eng = engine.Engine(local)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
bs = [eng.batch(n) for _ in range(8)]
for b in bs:
    b.encode(jobs)
    b.dispatch()
w0, w1 = shard.worker_range(rank, world, W)
host = torch.from_numpy(cfg.workers.loads()[w0:w1].view(np.uint8).reshape(-1, 16).copy()).pin_memory()
send = torch.empty((w1 - w0, 16), dtype=torch.uint8, device="cuda")
recv = [torch.empty((W, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
stream = torch.cuda.current_stream()


def step(k):
    send.copy_(host, non_blocking=True)
    buf = shard.gather_loads(send, out=recv[k % 2]) if world > 1 else send
    eng.set_loads_device(buf.data_ptr(), W, stream.cuda_stream)
    bs[k % 8].dispatch_resident_async()


def sync():
    for b in bs:
        b.wait()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for k in range(10):
    step(k)
sync()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for k in range(steps):
        step(k)
    sync()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "timeline_w%d.json" % world)
    prof.export_chrome_trace(path)
    tr = json.load(open(path))
    ev = [e for e in tr["traceEvents"] if e.get("ph") == "X"]
    gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    api = [e for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver")]
    t0 = min(e["ts"] for e in gpu)
    print("GPU activities (us from first):")
    for e in sorted(gpu, key=lambda e: e["ts"]):
        print("%9.1f %8.1f  s%-4s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:60]))
    print("runtime API calls: n=%d total %.1f us  (%.1f us/step)" % (len(api), sum(e["dur"] for e in api), sum(e["dur"] for e in api) / steps))
    from collections import defaultdict
    agg = defaultdict(lambda: [0, 0.0])
    for e in api:
        agg[e["name"]][0] += 1
        agg[e["name"]][1] += e["dur"]
    for k_, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("   %-32s n=%4d  %8.1f us  (%.2f us each)" % (k_, c, d, d / c))
    slim = {"traceEvents": gpu + api}
    json.dump(slim, open(path, "w"))
eng.close()
if world > 1:
    dist.destroy_process_group()
