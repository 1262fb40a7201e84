"""A runnable gRPC `SafetyKernel` service over the engine (grpcio + protobuf runtime; no protoc in the image, so the
descriptors are built programmatically).

    service SafetyKernel { rpc Check / Evaluate / Explain / Simulate (PolicyCheckRequest) returns (PolicyCheckResponse);
                           rpc ListSnapshots (ListSnapshotsRequest) returns (ListSnapshotsResponse); }
        = safetykernel.(*server).{Check,Evaluate,Explain,Simulate,ListSnapshots}   kernel.go:106-127

**Schema.**  Message and field NAMES and types restate what the reference uses of the CAP messages (SURVEY.md App. B:
`kernel.go:239-248,348-368,416-445`, `safety_client.go:80-95`).  The field NUMBERS are this file's own: the real .proto
lives in `github.com/cordum-io/cap/v2 v2.0.12`, which is not vendored in the reference tree, so they cannot be restated.
The service is therefore wire-compatible with clients built from this same descriptor (`SafetyKernelStub`), not with
CAP-generated clients; a Go deployment puts `go/cordumb200.SafetyKernel` behind the real generated service instead
(INTEGRATION.md).  What this module adds is the network-facing shape of the path: one blocking RPC per request on grpc's
thread pool, micro-batched onto the GPU by the C++ front-end (`cordum_frontend_*`), responses assembled from the policy
generation each request ran under.
"""
from __future__ import annotations

from concurrent import futures
from typing import Callable

import grpc
from google.protobuf import descriptor_pb2 as dp
from google.protobuf import descriptor_pool, message_factory

PACKAGE = "cordum.agent.v1"
SERVICE = PACKAGE + ".SafetyKernel"
_T = dp.FieldDescriptorProto


def _build_pool():
    f = dp.FileDescriptorProto(name="cordum_b200/safety_kernel.proto", package=PACKAGE, syntax="proto3")

    def enum(name, values):
        e = f.enum_type.add(name=name)
        for i, v in enumerate(values):
            e.value.add(name=v, number=i)

    def msg(name, fields):
        m = f.message_type.add(name=name)
        for num, (fname, ftype) in enumerate(fields, 1):
            rep = ftype.startswith("repeated ")
            ftype = ftype[9:] if rep else ftype
            if ftype.startswith("map<"):            # map<string,string>: the synthetic entry message protoc would emit
                ent = m.nested_type.add(name="".join(p.capitalize() for p in fname.split("_")) + "Entry")
                ent.options.map_entry = True
                ent.field.add(name="key", number=1, type=_T.TYPE_STRING, label=_T.LABEL_OPTIONAL)
                ent.field.add(name="value", number=2, type=_T.TYPE_STRING, label=_T.LABEL_OPTIONAL)
                m.field.add(name=fname, number=num, type=_T.TYPE_MESSAGE, label=_T.LABEL_REPEATED,
                            type_name=".%s.%s.%s" % (PACKAGE, name, ent.name))
                continue
            scalar = {"string": _T.TYPE_STRING, "bytes": _T.TYPE_BYTES, "bool": _T.TYPE_BOOL, "int32": _T.TYPE_INT32,
                      "int64": _T.TYPE_INT64, "double": _T.TYPE_DOUBLE}
            fd = m.field.add(name=fname, number=num, label=_T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL)
            if ftype in scalar:
                fd.type = scalar[ftype]
            elif ftype.startswith("enum "):
                fd.type, fd.type_name = _T.TYPE_ENUM, ".%s.%s" % (PACKAGE, ftype[5:])
            else:
                fd.type, fd.type_name = _T.TYPE_MESSAGE, ".%s.%s" % (PACKAGE, ftype)

    enum("ActorType", ["ACTOR_TYPE_UNSPECIFIED", "ACTOR_TYPE_HUMAN", "ACTOR_TYPE_SERVICE"])                     # kernel.go:370-379
    enum("DecisionType", ["DECISION_TYPE_UNSPECIFIED", "DECISION_TYPE_ALLOW", "DECISION_TYPE_DENY", "DECISION_TYPE_REQUIRE_HUMAN",
                          "DECISION_TYPE_THROTTLE", "DECISION_TYPE_ALLOW_WITH_CONSTRAINTS"])                    # pb.go:61-66
    msg("JobMetadata", [("tenant_id", "string"), ("actor_id", "string"), ("actor_type", "enum ActorType"), ("idempotency_key", "string"),
                        ("capability", "string"), ("risk_tags", "repeated string"), ("requires", "repeated string"), ("pack_id", "string"),
                        ("labels", "map<string,string>")])                                                      # kernel.go:357-363
    msg("PolicyCheckRequest", [("job_id", "string"), ("topic", "string"), ("tenant", "string"), ("principal_id", "string"),
                               ("labels", "map<string,string>"), ("meta", "JobMetadata"), ("effective_config", "bytes")])   # safety_client.go:80-95
    msg("BudgetConstraints", [("max_runtime_ms", "int64"), ("max_retries", "int32"), ("max_artifact_bytes", "int64"), ("max_concurrent_jobs", "int32")])
    msg("SandboxProfile", [("isolated", "bool"), ("network_allowlist", "repeated string"), ("fs_read_only", "repeated string"), ("fs_read_write", "repeated string")])
    msg("ToolchainConstraints", [("allowed_tools", "repeated string"), ("allowed_commands", "repeated string")])
    msg("DiffConstraints", [("max_files", "int32"), ("max_lines", "int32"), ("deny_path_globs", "repeated string")])
    msg("PolicyConstraints", [("budgets", "BudgetConstraints"), ("sandbox", "SandboxProfile"), ("toolchain", "ToolchainConstraints"),
                              ("diff", "DiffConstraints"), ("redaction_level", "string")])                      # kernel.go:416-445
    msg("PolicyRemediation", [("id", "string"), ("title", "string"), ("summary", "string"), ("replacement_topic", "string"),
                              ("replacement_capability", "string"), ("add_labels", "map<string,string>"), ("remove_labels", "repeated string")])   # kernel.go:328-346
    msg("PolicyCheckResponse", [("decision", "enum DecisionType"), ("reason", "string"), ("policy_snapshot", "string"), ("rule_id", "string"),
                                ("constraints", "PolicyConstraints"), ("approval_required", "bool"), ("approval_ref", "string"),
                                ("remediations", "repeated PolicyRemediation")])                                 # kernel.go:239-248
    msg("ListSnapshotsRequest", [])
    msg("ListSnapshotsResponse", [("snapshots", "repeated string")])                                            # kernel.go:122-127
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    return pool


_POOL = _build_pool()


def message(name: str):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(PACKAGE + "." + name))


PolicyCheckRequest, PolicyCheckResponse = message("PolicyCheckRequest"), message("PolicyCheckResponse")
ListSnapshotsRequest, ListSnapshotsResponse = message("ListSnapshotsRequest"), message("ListSnapshotsResponse")
DECISION_ENUM = {"ALLOW": 1, "DENY": 2, "REQUIRE_HUMAN": 3, "THROTTLE": 4, "ALLOW_WITH_CONSTRAINTS": 5}
DECISION_NAME = {v: k for k, v in DECISION_ENUM.items()}


def request_to_dict(req) -> dict:
    """PolicyCheckRequest -> the dict shape of reference_api / wire (a nil Meta stays absent: kernel.go:348-356)."""
    d = {"job_id": req.job_id, "topic": req.topic, "tenant": req.tenant, "principal_id": req.principal_id,
         "labels": dict(req.labels), "effective_config": bytes(req.effective_config)}
    if req.HasField("meta"):
        m = req.meta
        d["meta"] = {"tenant_id": m.tenant_id, "actor_id": m.actor_id, "actor_type": int(m.actor_type), "capability": m.capability,
                     "risk_tags": list(m.risk_tags), "requires": list(m.requires), "pack_id": m.pack_id}
    return d


def response_from_dict(r: dict):
    """reference_api.SafetyKernelServer's response dict -> PolicyCheckResponse (toProtoConstraints kernel.go:416-445,
    toProtoRemediations :328-346)."""
    out = PolicyCheckResponse(decision=DECISION_ENUM[r["decision"]], reason=r.get("reason", ""), policy_snapshot=r.get("policy_snapshot", ""),
                              rule_id=r.get("rule_id", ""), approval_required=bool(r.get("approval_required")), approval_ref=r.get("approval_ref", ""))
    c = r.get("constraints")
    if c:
        b, s, t, df = c.get("budgets") or {}, c.get("sandbox") or {}, c.get("toolchain") or {}, c.get("diff") or {}
        pc = out.constraints
        pc.budgets.max_runtime_ms = int(b.get("max_runtime_ms") or 0)
        pc.budgets.max_retries = int(b.get("max_retries") or 0)
        pc.budgets.max_artifact_bytes = int(b.get("max_artifact_bytes") or 0)
        pc.budgets.max_concurrent_jobs = int(b.get("max_concurrent_jobs") or 0)
        pc.sandbox.isolated = bool(s.get("isolated"))
        pc.sandbox.network_allowlist.extend(s.get("network_allowlist") or [])
        pc.sandbox.fs_read_only.extend(s.get("fs_read_only") or [])
        pc.sandbox.fs_read_write.extend(s.get("fs_read_write") or [])
        pc.toolchain.allowed_tools.extend(t.get("allowed_tools") or [])
        pc.toolchain.allowed_commands.extend(t.get("allowed_commands") or [])
        pc.diff.max_files = int(df.get("max_files") or 0)
        pc.diff.max_lines = int(df.get("max_lines") or 0)
        pc.diff.deny_path_globs.extend(df.get("deny_path_globs") or [])
        pc.redaction_level = c.get("redaction_level") or ""
    for rem in r.get("remediations") or []:
        pr = out.remediations.add(id=rem.get("id", ""), title=rem.get("title", ""), summary=rem.get("summary", ""),
                                  replacement_topic=rem.get("replacement_topic", ""), replacement_capability=rem.get("replacement_capability", ""))
        for k, v in (rem.get("add_labels") or {}).items():
            pr.add_labels[k] = v
        pr.remove_labels.extend(rem.get("remove_labels") or [])
    return out


def response_to_dict(resp) -> dict:
    """For clients / tests: the response as the plain dict reference_api uses."""
    d = {"decision": DECISION_NAME.get(resp.decision, "UNSPECIFIED"), "reason": resp.reason, "policy_snapshot": resp.policy_snapshot,
         "rule_id": resp.rule_id, "approval_required": resp.approval_required, "approval_ref": resp.approval_ref,
         "constraints": None, "remediations": []}
    if resp.HasField("constraints"):
        c = resp.constraints
        d["constraints"] = {"budgets": {"max_runtime_ms": c.budgets.max_runtime_ms, "max_retries": c.budgets.max_retries,
                                        "max_artifact_bytes": c.budgets.max_artifact_bytes, "max_concurrent_jobs": c.budgets.max_concurrent_jobs},
                            "redaction_level": c.redaction_level}
    for r in resp.remediations:
        d["remediations"].append({"id": r.id, "title": r.title, "replacement_topic": r.replacement_topic, "add_labels": dict(r.add_labels)})
    return d


class SafetyKernelServicer:
    """evaluate: dict -> dict (one request; must be thread-safe), list_snapshots: () -> [str].  All four RPCs are the
    same function (kernel.go:129: the mode string is ignored)."""

    def __init__(self, evaluate: Callable[[dict], dict], list_snapshots: Callable[[], list]):
        self._evaluate, self._list = evaluate, list_snapshots

    def _check(self, request, context):
        try:
            return response_from_dict(self._evaluate(request_to_dict(request)))
        except Exception as exc:   # evaluate never returns a gRPC error for policy outcomes; an engine failure fails closed
            return PolicyCheckResponse(decision=DECISION_ENUM["DENY"], reason="safety kernel error: %s" % exc)

    def _snapshots(self, request, context):
        return ListSnapshotsResponse(snapshots=list(self._list()))

    def handler(self):
        unary = lambda fn, req_cls, resp_cls: grpc.unary_unary_rpc_method_handler(   # noqa: E731
            fn, request_deserializer=req_cls.FromString, response_serializer=resp_cls.SerializeToString)
        methods = {name: unary(self._check, PolicyCheckRequest, PolicyCheckResponse) for name in ("Check", "Evaluate", "Explain", "Simulate")}
        methods["ListSnapshots"] = unary(self._snapshots, ListSnapshotsRequest, ListSnapshotsResponse)
        return grpc.method_handlers_generic_handler(SERVICE, methods)


class SafetyKernelStub:
    """Client side of the same descriptor."""

    def __init__(self, channel):
        mk = lambda name, req, resp: channel.unary_unary("/%s/%s" % (SERVICE, name), request_serializer=req.SerializeToString,   # noqa: E731
                                                         response_deserializer=resp.FromString)
        self.Check, self.Evaluate = mk("Check", PolicyCheckRequest, PolicyCheckResponse), mk("Evaluate", PolicyCheckRequest, PolicyCheckResponse)
        self.Explain, self.Simulate = mk("Explain", PolicyCheckRequest, PolicyCheckResponse), mk("Simulate", PolicyCheckRequest, PolicyCheckResponse)
        self.ListSnapshots = mk("ListSnapshots", ListSnapshotsRequest, ListSnapshotsResponse)


def request_from_dict(job: dict):
    req = PolicyCheckRequest(job_id=job.get("job_id", ""), topic=job.get("topic", ""), tenant=job.get("tenant", ""),
                             principal_id=job.get("principal_id", ""), effective_config=bytes(job.get("effective_config") or b""))
    for k, v in (job.get("labels") or {}).items():
        req.labels[k] = v
    meta = job.get("meta")
    if meta is not None:
        at = meta.get("actor_type", 0)
        if isinstance(at, str):
            at = {"human": 1, "service": 2}.get(at.lower(), 0)
        req.meta.SetInParent()
        req.meta.tenant_id, req.meta.actor_id, req.meta.actor_type = meta.get("tenant_id", ""), meta.get("actor_id", ""), at
        req.meta.capability, req.meta.pack_id = meta.get("capability", ""), meta.get("pack_id", "")
        req.meta.risk_tags.extend(meta.get("risk_tags") or [])
        req.meta.requires.extend(meta.get("requires") or [])
    return req


def serve(servicer: SafetyKernelServicer, address: str = "127.0.0.1:0", max_workers: int = 64):
    """-> (grpc.Server, bound port).  insecure transport, as the reference without SAFETY_KERNEL_TLS_CERT (kernel.go:75-85)."""
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    server.add_generic_rpc_handlers((servicer.handler(),))
    port = server.add_insecure_port(address)
    server.start()
    return server, port


def engine_servicer(engine, cache_ttl_us: int = 0, max_batch: int = 1024, max_wait_us: int = 200) -> SafetyKernelServicer:
    """The GPU engine behind the service: every RPC thread makes one blocking `cordum_frontend_submit`; the C++ front-end
    batches whatever is in flight (ctypes releases the interpreter lock for the call)."""
    from . import frontend, wire

    fe = frontend.Frontend(engine, max_batch=max_batch, max_wait_us=max_wait_us, mode=wire.MODE_POLICY_ONLY, cache_ttl_us=cache_ttl_us)
    names = {wire.DEC_ALLOW: "ALLOW", wire.DEC_DENY: "DENY", wire.DEC_REQUIRE_HUMAN: "REQUIRE_HUMAN", wire.DEC_THROTTLE: "THROTTLE",
             wire.DEC_ALLOW_WITH_CONSTRAINTS: "ALLOW_WITH_CONSTRAINTS"}

    def evaluate(req: dict) -> dict:
        r = fe.submit(req)
        if r.status != 0:
            raise RuntimeError(r.reason.decode("utf-8", "replace"))
        flags, idx, gen = int(r.rec.flags), int(r.rec.rule_idx), int(r.policy_gen)
        approval = bool(flags & wire.F_APPROVAL_REQUIRED)
        return {"decision": names[int(r.rec.decision)], "reason": r.reason.decode("utf-8", "replace"),
                "policy_snapshot": r.snapshot.decode("utf-8", "replace"), "rule_id": r.rule_id.decode("utf-8", "replace"),
                "constraints": engine.rule_constraints(idx, gen) if flags & wire.F_CONSTRAINTS else None,
                "approval_required": approval, "approval_ref": req.get("job_id", "") if approval else "",     # kernel.go:233-237
                "remediations": engine.rule_remediations(idx, gen) if idx >= 0 else []}

    sv = SafetyKernelServicer(evaluate, engine.snapshots)
    sv.frontend = fe
    return sv
