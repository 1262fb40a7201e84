"""Where a sharded bench step spends its time.  Run under torchrun:
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scalebench.py [jobs_per_rank]
Variants (all: K steps, wall clock between synchronizes, max over ranks):
   dispatch      resident dispatch only (no heartbeat epoch)
   ingest        H2D slice + set_loads_device (no collective) + dispatch
   allgather     the collective alone
   full          H2D slice + all-gather + set_loads_device + dispatch on one ingest stream
   full2         the same with two alternating ingest streams and buffers (= bench.py's step)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

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
n = int(sys.argv[1]) if len(sys.argv) > 1 else (J + world - 1) // world
j0 = (rank * n) % max(1, J - n)
jobs = cfg.jobs.slice(j0, n)
eng = engine.Engine(local)
eng.load_policy(cfg.policy, "x")
eng.load_routing(cfg.routing)
eng.load_workers(cfg.workers)
n_rot = 8
bs = [eng.batch(n) for _ in range(n_rot)]
for b in bs:
    b.encode(jobs)
    b.dispatch()
w0, w1 = shard.worker_range(rank, world, W)
host = torch.from_numpy(cfg.workers.loads()[w0:w1].view(np.uint8).reshape(-1, 16).copy()).pin_memory()
full_host = torch.from_numpy(cfg.workers.loads().view(np.uint8).reshape(-1, 16).copy()).pin_memory()
send = torch.empty((w1 - w0, 16), dtype=torch.uint8, device="cuda")
send2 = [torch.empty((w1 - w0, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
ingest = [torch.cuda.Stream(), torch.cuda.Stream()]
full_dev = torch.empty((W, 16), dtype=torch.uint8, device="cuda")
full_dev.copy_(full_host)
recv = [torch.empty((W, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
stream = torch.cuda.current_stream()


def sync():
    for b in bs:
        b.wait()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(variant, K):
    th = 0.0
    for k in range(K):
        a = time.perf_counter()
        if variant in ("full", "allgather"):
            send.copy_(host, non_blocking=True)
            buf = shard.gather_loads(send, out=recv[k % 2]) if world > 1 else send
            if variant == "full":
                eng.set_loads_device(buf.data_ptr(), W, stream.cuda_stream)
        elif variant == "full2":
            with torch.cuda.stream(ingest[k % 2]):
                send2[k % 2].copy_(host, non_blocking=True)
                buf = shard.gather_loads(send2[k % 2], out=recv[k % 2]) if world > 1 else send2[k % 2]
                eng.set_loads_device(buf.data_ptr(), W, ingest[k % 2].cuda_stream)
        elif variant == "ingest":
            send.copy_(host, non_blocking=True)
            eng.set_loads_device(full_dev.data_ptr(), W, stream.cuda_stream)
        if variant != "allgather":
            bs[k % n_rot].dispatch_resident_async()
        th += time.perf_counter() - a
    return th


for variant in ("dispatch", "ingest", "allgather", "full", "full2", "full", "full2"):
    run(variant, 10)
    sync()
    K = 200
    t0 = time.perf_counter()
    th = run(variant, K)
    t_enq = time.perf_counter() - t0
    for b in bs:
        b.wait()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, th, t_enq], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, th, t_enq = (float(x) for x in t.tolist())
    if rank == 0:
        print("%-10s world %d  %d jobs/rank  step %.1f us  host enqueue %.1f us  kernels %s" % (
            variant, world, n, dt / K * 1e6, th / K * 1e6, tuple(round(x * 1e3, 1) for x in bs[0].kernel_times())), flush=True)
    sync()
eng.close()
if world > 1:
    dist.destroy_process_group()
