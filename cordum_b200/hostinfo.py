"""Host facts the bench and tests report: how many cores this process may actually use."""
from __future__ import annotations

import math
import os


def cgroup_cpu_quota() -> float | None:
    """CPU quota of the container in cores (cgroup v2 cpu.max, else v1 cfs quota), None = unlimited."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()[:2]
        if q != "max" and float(per) > 0:
            return float(q) / float(per)
        return None
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
            q = float(f.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            per = float(f.read())
        if q > 0 and per > 0:
            return q / per
    except (OSError, ValueError):
        pass
    return None


def usable_cores() -> int:
    """min(scheduler affinity, cgroup quota): the number of threads that can run flat out (os.cpu_count() reports the
    machine, not the container: 128 on boxes whose quota is 24)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    q = cgroup_cpu_quota()
    if q is not None and q >= 1.0:
        n = min(n, int(math.floor(q + 0.5)))
    return max(1, n)


def describe() -> dict:
    try:
        aff = len(os.sched_getaffinity(0))
    except AttributeError:
        aff = None
    return {"os_cpu_count": os.cpu_count(), "affinity": aff, "cgroup_quota": cgroup_cpu_quota(), "usable": usable_cores()}
