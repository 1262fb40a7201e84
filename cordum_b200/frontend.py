"""ctypes view of the micro-batching front-end (include/cordum_b200.h cordum_frontend_*): one blocking call per request,
any number of threads; concurrent callers are served by one encode + dispatch per batch."""
from __future__ import annotations

import ctypes as C

from . import wire


class Sv(C.Structure):
    _fields_ = [("p", C.c_char_p), ("n", C.c_uint32)]


class Kv(C.Structure):
    _fields_ = [("key", Sv), ("val", Sv)]


class Request(C.Structure):
    _fields_ = [("topic", Sv), ("tenant", Sv), ("principal_id", Sv), ("effective_config", Sv),
                ("has_meta", C.c_uint8), ("actor_type", C.c_uint8), ("approved", C.c_uint8), ("pad", C.c_uint8),
                ("meta_tenant_id", Sv), ("actor_id", Sv), ("capability", Sv), ("pack_id", Sv),
                ("risk_tags", C.POINTER(Sv)), ("n_risk_tags", C.c_uint32),
                ("requires_", C.POINTER(Sv)), ("n_requires", C.c_uint32),
                ("labels", C.POINTER(Kv)), ("n_labels", C.c_uint32)]


class Decision(C.Structure):
    _fields_ = [("decision", C.c_uint8), ("sched_decision", C.c_uint8), ("flags", C.c_uint8), ("route_status", C.c_uint8),
                ("reason_code", C.c_uint8), ("reserved", C.c_uint8 * 3), ("rule_idx", C.c_int32), ("worker_slot", C.c_int32)]


class Response(C.Structure):
    _fields_ = [("rec", Decision), ("status", C.c_int32), ("reserved", C.c_uint32), ("policy_gen", C.c_uint64), ("rule_id", C.c_char * 128), ("reason", C.c_char * 256),
                ("subject", C.c_char * 192), ("snapshot", C.c_char * 96)]


class FrontendOpts(C.Structure):
    _fields_ = [("max_batch", C.c_uint32), ("max_wait_us", C.c_uint32), ("mode", C.c_uint32), ("lanes", C.c_uint32),
                ("arena_bytes_per_request", C.c_uint32), ("reserved", C.c_uint32), ("cache_ttl_us", C.c_uint64)]


def _b(s) -> bytes:
    if s is None:
        return b""
    return s if isinstance(s, (bytes, bytearray)) else str(s).encode("utf-8", "surrogatepass")


def pack_request(job: dict):
    """(Request, keepalive) for a job dict in the shape of cordum_b200/wire.py."""
    keep = []

    def sv(x):
        b = _b(x)
        keep.append(b)
        return Sv(b if b else None, len(b))

    r = Request()
    r.topic, r.tenant, r.principal_id = sv(job.get("topic", "")), sv(job.get("tenant", "")), sv(job.get("principal_id", ""))
    r.effective_config = sv(job.get("effective_config") or b"")
    meta = job.get("meta")
    r.approved = 1 if job.get("approved") else 0
    if meta is not None:
        r.has_meta = 1
        at = meta.get("actor_type", 0)
        if isinstance(at, str):
            at = {"human": 1, "service": 2}.get(at.lower(), 0)
        r.actor_type = at
        r.meta_tenant_id, r.actor_id = sv(meta.get("tenant_id", "")), sv(meta.get("actor_id", ""))
        r.capability, r.pack_id = sv(meta.get("capability", "")), sv(meta.get("pack_id", ""))
        tags = [sv(t) for t in (meta.get("risk_tags") or [])]
        reqs = [sv(t) for t in (meta.get("requires") or [])]
        if tags:
            arr = (Sv * len(tags))(*tags)
            keep.append(arr)
            r.risk_tags, r.n_risk_tags = arr, len(tags)
        if reqs:
            arr = (Sv * len(reqs))(*reqs)
            keep.append(arr)
            r.requires_, r.n_requires = arr, len(reqs)
    labels = list((job.get("labels") or {}).items())
    if labels:
        arr = (Kv * len(labels))(*[Kv(sv(k), sv(v)) for k, v in labels])
        keep.append(arr)
        r.labels, r.n_labels = arr, len(labels)
    return r, keep


class Frontend:
    def __init__(self, eng, max_batch=1024, max_wait_us=200, mode=wire.MODE_POLICY_AND_ROUTE, lanes=2, arena_bytes_per_request=1024,
                 cache_ttl_us=0):
        """cache_ttl_us: SAFETY_DECISION_CACHE_TTL (kernel.go:149-162); only POLICY_ONLY front-ends cache."""
        self.eng = eng
        self.L = eng.L
        opts = FrontendOpts(max_batch, max_wait_us, mode, lanes, arena_bytes_per_request, 0, cache_ttl_us)
        h = C.c_void_p()
        eng._ck(self.L.cordum_frontend_create(eng.h, C.byref(opts), C.byref(h)))
        self.h = h

    def submit(self, job) -> Response:
        """Blocking; call from as many threads as you like (ctypes releases the GIL for the duration of the call)."""
        req, keep = job if isinstance(job, tuple) else pack_request(job)
        resp = Response()
        self.L.cordum_frontend_submit(self.h, C.byref(req), C.byref(resp))
        return resp

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.cordum_frontend_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"batches": a.value, "requests": b.value, "full_batches": c.value}

    def cache_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.cordum_frontend_cache_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"hits": a.value, "misses": b.value, "entries": c.value}

    def close(self):
        if self.h:
            self.L.cordum_frontend_destroy(self.h)
            self.h = None
