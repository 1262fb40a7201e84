"""ctypes access to oracle/liboracle.so — the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use this.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

from cordum_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.cpp", "oracle.h")] + [
        os.path.join(ROOT, "include", "cordum_b200.h"), os.path.join(ROOT, "common", "mini_json.hpp"),
        os.path.join(ROOT, "common", "go_unicode_tables.h")]
    import fcntl

    def stale():
        return force or not os.path.exists(LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)

    if stale():
        # several processes (one per GPU under torchrun, pytest-xdist workers) may get here together: one builds, the
        # others wait for the lock and then find the library fresh
        with open(os.path.join(ORACLE_DIR, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if stale():
                    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_policy_load.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.oracle_routing_load.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.oracle_workers_load.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_workers_update.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_eval_one_json.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64]
        L.oracle_eval_one_json.restype = C.c_int64
        L.oracle_eval_one_json_flavor.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64]
        L.oracle_eval_one_json_flavor.restype = C.c_int64
        L.oracle_quote.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
        L.oracle_quote.restype = C.c_int64
        L.oracle_path_match.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
        L.oracle_equal_fold.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
        L.oracle_to_lower.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
        L.oracle_to_lower.restype = C.c_int64
        L.oracle_trim_space.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_normalize_decision.argtypes = [C.c_char_p, C.c_uint64]
        L.oracle_parse_effective.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else str(s).encode("utf-8", "surrogatepass")


def path_match(pat, name) -> int:
    p, n = _b(pat), _b(name)
    return lib().oracle_path_match(p, len(p), n, len(n))


def equal_fold(a, b) -> bool:
    a, b = _b(a), _b(b)
    return bool(lib().oracle_equal_fold(a, len(a), b, len(b)))


def to_lower(s) -> bytes:
    s = _b(s)
    buf = C.create_string_buffer(4 * len(s) + 8)
    n = lib().oracle_to_lower(s, len(s), buf, len(buf))
    return buf.raw[:n]


def quote(s) -> bytes:
    """strconv.Quote"""
    s = _b(s)
    buf = C.create_string_buffer(10 * len(s) + 8)
    n = lib().oracle_quote(s, len(s), buf, len(buf))
    return buf.raw[:n]


def trim_space(s) -> bytes:
    s = _b(s)
    off, ln = C.c_uint64(), C.c_uint64()
    lib().oracle_trim_space(s, len(s), C.byref(off), C.byref(ln))
    return s[off.value: off.value + ln.value]


def normalize_decision(s) -> int:
    s = _b(s)
    return lib().oracle_normalize_decision(s, len(s))


def parse_effective(s):
    s = _b(s)
    a, d = C.c_uint32(), C.c_uint32()
    ok = lib().oracle_parse_effective(s, len(s), C.byref(a), C.byref(d))
    return bool(ok), a.value, d.value


class Oracle:
    """One oracle context: policy + routing + worker registry."""

    def __init__(self, policy=None, routing=None, workers=None):
        self.L = lib()
        self.h = C.c_void_p(self.L.oracle_create())
        self._wt = None
        self.load_policy(policy)
        self.load_routing(routing)
        self.load_workers(workers or [])

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: %s" % self.L.oracle_last_error().decode())

    def load_policy(self, policy):
        doc = b"" if policy is None else (policy if isinstance(policy, bytes) else json.dumps(policy).encode())
        self._check(self.L.oracle_policy_load(self.h, doc, len(doc)))

    def load_routing(self, routing):
        doc = b"" if routing is None else (routing if isinstance(routing, bytes) else json.dumps(routing).encode())
        self._check(self.L.oracle_routing_load(self.h, doc, len(doc)))

    def load_workers(self, workers):
        wt = workers if isinstance(workers, wire.WorkerTable) else wire.WorkerTable.from_workers(workers)
        self._wt = wt
        self._check(self.L.oracle_workers_load(self.h, C.addressof(wt.struct)))

    def update_workers(self, slots, loads):
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        loads = np.ascontiguousarray(loads, dtype=wire.LOAD_DTYPE)
        self._check(self.L.oracle_workers_update(self.h, len(slots), slots.ctypes.data, loads.ctypes.data))

    def eval(self, env, mode=wire.MODE_POLICY_AND_ROUTE, threads=1, first=0, count=None) -> np.ndarray:
        if not isinstance(env, wire.EnvelopeBatch):
            env = wire.EnvelopeBatch.from_jobs(env)
        if count is None:
            count = env.n_jobs - first
        out = np.zeros(max(count, 1), dtype=wire.DECISION_DTYPE)
        self._check(self.L.oracle_eval(self.h, C.addressof(env.struct), first, count, mode, threads, out.ctypes.data))
        return out[:count]

    def eval_one(self, job, mode=wire.MODE_POLICY_AND_ROUTE, flavor=0) -> dict:
        """flavor 1 = the gateway's evaluatePolicyCheck (policy_bundles.go:1132-1231)"""
        env = job if isinstance(job, wire.EnvelopeBatch) else wire.EnvelopeBatch.from_jobs([job])
        buf = C.create_string_buffer(1 << 16)
        n = self.L.oracle_eval_one_json_flavor(self.h, C.addressof(env.struct), 0, mode, flavor, buf, len(buf))
        assert 0 <= n < len(buf)
        return json.loads(buf.value.decode("utf-8", "replace"))
