"""py_oracle.py — second, independent CPU restatement of the reference path.

*** TEST INFRASTRUCTURE. NOT PRODUCT CODE. ***  Pure Python, dict-level, slow on
purpose: it exists to cross-check oracle/oracle.cpp on the reference's own
known-answer tests and on randomized inputs ("two independent restatements",
SURVEY.md §7 item 1).  Only tests/ may import it.

Each function cites the Go it follows (paths relative to /root/reference):
  core/infra/config/safety_policy.go          -> Policy.*
  core/controlplane/safetykernel/kernel.go    -> kernel_evaluate and helpers
  core/infra/config/effective.go              -> parse_effective_safety
  core/controlplane/scheduler/strategy_least_loaded.go -> pick_subject and helpers
  core/controlplane/scheduler/engine.go:298-347,484-531 -> process_job
Go stdlib restated: strings.TrimSpace / EqualFold (Unicode simple case folding) / ToLower
(simple lowercase mapping), path.Match (go1.24 src/path/match.go).

Inputs: policy = dict shaped like config.SafetyPolicy with yaml tag names (or None);
routing = {"topics": {t: [pools]}, "pools": {p: {"requires": [...]}}};
workers = list of heartbeat dicts; job = dict as in cordum_b200/wire.py.
"""
from __future__ import annotations

import json
import struct
import unicodedata

# ------------------------------------------------------------------ Go strings
_GO_SPACE = {chr(c) for c in (0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000)} | {chr(c) for c in range(0x2000, 0x200B)}


def trim_space(s: str) -> str:
    a, b = 0, len(s)
    while a < b and s[a] in _GO_SPACE:
        a += 1
    while b > a and s[b - 1] in _GO_SPACE:
        b -= 1
    return s[a:b]


def _fold_rep(c: str) -> str:
    """Representative of c's class under Go's rune-level EqualFold (simple case folding, Unicode 15.0.0 = this
    interpreter's unicodedata).  Derived independently of common/go_unicode_tables.h from Python's own case mappings:
    a single-character full case folding is the simple folding; where the full folding expands (U+1E9E, the Greek
    iota-subscript capitals ...) the simple folding is the single-character lower case; characters whose only foldings
    are multi-character or Turkic (U+00DF, U+0130, U+0131, U+0390 ...) fold with nothing but their own lower/upper pair."""
    f = c.casefold()
    if len(f) == 1:
        return f
    low = c.lower()
    return low if len(low) == 1 else c


def equal_fold(a: str, b: str) -> bool:
    """strings.EqualFold: rune by rune, two runes are equal iff they are in the same simple-folding class."""
    return len(a) == len(b) and all(x == y or _fold_rep(x) == _fold_rep(y) for x, y in zip(a, b))


def to_lower(s: str) -> str:
    """strings.ToLower: unicode.ToLower per rune = the simple lowercase mapping (U+0130 -> 'i', not 'i' + U+0307)."""
    import _sre

    return "".join(chr(_sre.unicode_tolower(ord(c))) for c in s)


class BadPattern(Exception):
    pass


def _scan_chunk(pattern: str):
    star = False
    while pattern and pattern[0] == "*":
        pattern = pattern[1:]
        star = True
    inrange = False
    i = 0
    n = len(pattern)
    while i < n:
        c = pattern[i]
        if c == "\\":
            if i + 1 < n:
                i += 1
        elif c == "[":
            inrange = True
        elif c == "]":
            inrange = False
        elif c == "*":
            if not inrange:
                break
        i += 1
    return star, pattern[:i], pattern[i:]


def _get_esc(chunk: str):
    if not chunk or chunk[0] in "-]":
        raise BadPattern
    if chunk[0] == "\\":
        chunk = chunk[1:]
        if not chunk:
            raise BadPattern
    r = chunk[0]
    chunk = chunk[1:]
    if not chunk:
        raise BadPattern
    return r, chunk


def _match_chunk(chunk: str, s: str):
    """returns (rest, ok); raises BadPattern."""
    failed = False
    while chunk:
        if not failed and not s:
            failed = True
        c = chunk[0]
        if c == "[":
            r = ""
            if not failed:
                r, s = s[0], s[1:]
            chunk = chunk[1:]
            negated = False
            if chunk and chunk[0] == "^":
                negated = True
                chunk = chunk[1:]
            match = False
            nrange = 0
            while True:
                if chunk and chunk[0] == "]" and nrange > 0:
                    chunk = chunk[1:]
                    break
                lo, chunk = _get_esc(chunk)
                hi = lo
                if chunk[0] == "-":
                    hi, chunk = _get_esc(chunk[1:])
                if not failed and lo <= r <= hi:
                    match = True
                nrange += 1
            if match == negated:
                failed = True
        elif c == "?":
            if not failed:
                if s[0] == "/":
                    failed = True
                s = s[1:]
            chunk = chunk[1:]
        else:
            if c == "\\":
                chunk = chunk[1:]
                if not chunk:
                    raise BadPattern
            if not failed:
                # Go compares bytes; for str inputs code-point compare is equivalent
                if chunk[0] != s[0]:
                    failed = True
                s = s[1:]
            chunk = chunk[1:]
    if failed:
        return "", False
    return s, True


def path_match(pattern: str, name: str) -> bool:
    """path.Match; raises BadPattern."""
    while pattern:
        star, chunk, pattern = _scan_chunk(pattern)
        if star and chunk == "":
            return "/" not in name
        t, ok = _match_chunk(chunk, name)
        if ok and (len(t) == 0 or len(pattern) > 0):
            name = t
            continue
        advanced = False
        if star:
            i = 0
            while i < len(name) and name[i] != "/":
                t, ok = _match_chunk(chunk, name[i + 1:])
                if ok:
                    if len(pattern) == 0 and len(t) > 0:
                        i += 1
                        continue
                    name = t
                    advanced = True
                    break
                i += 1
        if advanced:
            continue
        while pattern:
            _, chunk, pattern = _scan_chunk(pattern)
            _match_chunk(chunk, "")
        return False
    return len(name) == 0


def glob_ok(pattern: str, value: str) -> bool:
    """matchTopic (safety_policy.go:356-363) / configMatch (kernel.go:467-474)."""
    pattern = trim_space(pattern)
    if pattern == "":
        return False
    try:
        return path_match(pattern, value)
    except BadPattern:
        return False


# ------------------------------------------------------------------ policy
def normalize_decision(raw) -> str:   # safety_policy.go:208-223
    s = to_lower(trim_space(raw or ""))
    if s in ("allow", "permit"):
        return "allow"
    if s in ("deny", "block"):
        return "deny"
    if s in ("require_approval", "require-approval", "require_human"):
        return "require_approval"
    if s in ("allow_with_constraints", "allow-with-constraints"):
        return "allow_with_constraints"
    if s == "throttle":
        return "throttle"
    return "allow"


def contains_string(lst, value) -> bool:   # :296-306
    if value == "":
        return False
    tv = trim_space(value)
    return any(equal_fold(trim_space(v or ""), tv) for v in (lst or []))


def contains_any(lst, values) -> bool:   # :308-318
    if not lst or not values:
        return False
    return any(contains_string(lst, v) for v in values)


def contains_all(values, required) -> bool:   # :320-330
    return all(contains_string(values, v or "") for v in (required or []))


def labels_match(required, actual) -> bool:   # :332-345
    if not required:
        return True
    if not actual:
        return False
    return all(actual.get(k, "") == (v or "") for k, v in required.items())


MCP_FIELDS = ("server", "tool", "resource", "action")
_MCP_KEYS = {"server": ("allow_servers", "deny_servers"), "tool": ("allow_tools", "deny_tools"),
             "resource": ("allow_resources", "deny_resources"), "action": ("allow_actions", "deny_actions")}


def mcp_used(req) -> bool:   # :404-406
    return any(trim_space(req[f]) != "" for f in MCP_FIELDS)


_C_ESCAPES = {7: "\\a", 8: "\\b", 12: "\\f", 10: "\\n", 13: "\\r", 9: "\\t", 11: "\\v"}


def go_quote(s) -> str:
    """fmt %q on a string = strconv.Quote.  IsPrint (letters, marks, numbers, punctuation, symbols, ASCII space) comes
    from Python's own unicodedata (15.0.0 in this image's CPython 3.12, the version go1.24 uses), not from the generated
    table the C++ sides share.  A str is taken as valid UTF-8 (lone surrogates = undecodable bytes via surrogateescape);
    bytes are decoded the way Go ranges over a string."""
    if isinstance(s, (bytes, bytearray)):
        s = bytes(s).decode("utf-8", "surrogateescape")
    out = ['"']
    for ch in s:
        r = ord(ch)
        if 0xDC80 <= r <= 0xDCFF:                       # an undecodable byte
            out.append("\\x%02x" % (r - 0xDC00))
        elif ch in '"\\':
            out.append("\\" + ch)
        elif r == 0x20 or unicodedata.category(ch)[0] in "LMNPS":
            out.append(ch)
        elif r in _C_ESCAPES:
            out.append(_C_ESCAPES[r])
        elif r < 0x20 or r == 0x7F:
            out.append("\\x%02x" % r)
        elif 0xD800 <= r <= 0xDFFF:                     # cannot come out of a Go string; RuneError
            out.append("\\ufffd")
        elif r < 0x10000:
            out.append("\\u%04x" % r)
        else:
            out.append("\\U%08x" % r)
    out.append('"')
    return "".join(out)


def mcp_allowed(policy, req):   # :385-402 ; returns (ok, reason, code) code = field*2 + notallowed
    policy = policy or {}
    if not mcp_used(req):
        return True, "", -1
    for fi, f in enumerate(MCP_FIELDS):
        allow = policy.get(_MCP_KEYS[f][0]) or []
        deny = policy.get(_MCP_KEYS[f][1]) or []
        v = req[f]
        if contains_string(deny, v):
            return False, "mcp %s %s denied" % (f, go_quote(v)), fi * 2
        if len(allow) > 0 and not contains_string(allow, v):
            return False, "mcp %s %s not allowed" % (f, go_quote(v)), fi * 2 + 1
    return True, "", -1


def match_rule(m, inp) -> bool:   # :259-294
    m = m or {}
    if m.get("tenants") and not contains_string(m["tenants"], inp["tenant"]):
        return False
    if m.get("topics") and not any(glob_ok(p or "", inp["topic"]) for p in m["topics"]):
        return False
    meta = inp["meta"]
    if m.get("capabilities") and not contains_string(m["capabilities"], meta["capability"]):
        return False
    if m.get("risk_tags") and not contains_any(m["risk_tags"], meta["risk_tags"]):
        return False
    if m.get("requires") and not contains_all(meta["requires"], m["requires"]):
        return False
    if m.get("pack_ids") and not contains_string(m["pack_ids"], meta["pack_id"]):
        return False
    if m.get("actor_ids") and not contains_string(m["actor_ids"], meta["actor_id"]):
        return False
    if m.get("actor_types") and not contains_string(m["actor_types"], meta["actor_type"]):
        return False
    sp = m.get("secrets_present")
    if sp is not None and bool(inp["secrets_present"]) != bool(sp):
        return False
    if m.get("labels") and not labels_match(m["labels"], inp["labels"]):
        return False
    ok, _, _ = mcp_allowed(m.get("mcp"), inp["mcp"])
    return ok


def effective_rules(policy):   # :188-191, 225-257 (tenants in sorted order)
    rules = policy.get("rules") or []
    if rules:
        return rules
    out = []
    for tenant in sorted((policy.get("tenants") or {}).keys(), key=lambda s: s.encode()):
        tp = policy["tenants"][tenant] or {}
        for i, pat in enumerate(tp.get("deny_topics") or []):
            out.append({"id": "legacy:%s:deny:%d" % (tenant, i + 1), "decision": "deny",
                        "reason": "topic %s denied by tenant policy" % go_quote(pat),
                        "match": {"tenants": [tenant], "topics": [pat], "mcp": tp.get("mcp")}})
        for i, pat in enumerate(tp.get("allow_topics") or []):
            out.append({"id": "legacy:%s:allow:%d" % (tenant, i + 1), "decision": "allow", "reason": "",
                        "match": {"tenants": [tenant], "topics": [pat], "mcp": tp.get("mcp")}})
    return out


def policy_evaluate(policy, inp):   # :187-206
    for idx, rule in enumerate(effective_rules(policy)):
        if match_rule(rule.get("match"), inp):
            d = normalize_decision(rule.get("decision"))
            return {"decision": d, "reason": rule.get("reason") or "", "rule_id": rule.get("id") or "",
                    "rule_idx": idx, "approval_required": d == "require_approval",
                    "constraints": rule.get("constraints") or {}}
    return {"decision": "allow", "reason": "", "rule_id": "", "rule_idx": -1, "approval_required": False,
            "constraints": {}}


def constraints_empty(c) -> bool:   # kernel.go:447-453
    c = c or {}
    b, s, t, d = (c.get("budgets") or {}), (c.get("sandbox") or {}), (c.get("toolchain") or {}), (c.get("diff") or {})
    return (not b.get("max_runtime_ms") and not b.get("max_retries") and not b.get("max_artifact_bytes")
            and not b.get("max_concurrent_jobs") and not s.get("isolated") and not s.get("network_allowlist")
            and not s.get("fs_read_only") and not s.get("fs_read_write") and not t.get("allowed_tools")
            and not t.get("allowed_commands") and not d.get("max_files") and not d.get("max_lines")
            and not d.get("deny_path_globs") and trim_space(c.get("redaction_level") or "") == "")


# ------------------------------------------------------------------ effective config
_SAFETY_FIELDS = {
    "pii_detection_enabled": "b", "pii_action": "s", "pii_types": "l", "allowed_email_domains": "l",
    "injection_detection": "b", "injection_action": "s", "injection_sensitivity": "s",
    "content_filter_enabled": "b", "blocked_categories": "l", "anomaly_detection": "b",
    "anomaly_thresholds": "m", "allowed_topics": "l", "denied_topics": "l", "allowed_repo_hosts": "l",
    "denied_repo_hosts": "l", "mcp": "p"}
_MCP_LISTS = ("allow_servers", "deny_servers", "allow_tools", "deny_tools", "allow_resources", "deny_resources",
              "allow_actions", "deny_actions")


def _field(key, names):
    if key in names:
        return key
    for n in names:
        if equal_fold(key, n):
            return n
    return None


def _is_num(x):   # a JSON number that fits float64 (strconv.ParseFloat reports a range error otherwise)
    if isinstance(x, bool) or not isinstance(x, (int, float)):
        return False
    try:
        return float(x) not in (float("inf"), float("-inf"))
    except OverflowError:
        return False


def _str_list(x):
    """([]string value, ok)"""
    if x is None:
        return [], True
    if not isinstance(x, list):
        return None, False
    ok = True
    out = []
    for e in x:
        if e is None:
            out.append("")
        elif isinstance(e, str):
            out.append(e)
        else:
            out.append("")
            ok = False
    return out, ok


def _decode_safety(pairs):
    """pairs: list of (key, value) in document order, or None for JSON null. -> (cfg, ok)"""
    cfg = {"allowed_topics": [], "denied_topics": [], "mcp": {k: [] for k in _MCP_LISTS}}
    if pairs is None:
        return cfg, True
    if not isinstance(pairs, _Obj):
        return cfg, False
    ok = True
    for key, x in pairs.items:
        f = _field(key, _SAFETY_FIELDS)
        if f is None:
            continue
        kind = _SAFETY_FIELDS[f]
        if kind == "b":
            ok &= x is None or isinstance(x, bool)
        elif kind == "s":
            ok &= x is None or isinstance(x, str)
        elif kind == "l":
            v, good = _str_list(x)
            ok &= good
            if good or v is not None:
                if f in ("allowed_topics", "denied_topics") and v is not None:
                    cfg[f] = v
        elif kind == "m":
            if x is None:
                pass
            elif not isinstance(x, _Obj):
                ok = False
            else:
                ok &= all(v is None or _is_num(v) for _, v in x.items)
        elif kind == "p":
            if x is None:
                pass
            elif not isinstance(x, _Obj):
                ok = False
            else:
                for k2, v2 in x.items:
                    f2 = _field(k2, _MCP_LISTS)
                    if f2 is None:
                        continue
                    v, good = _str_list(v2)
                    ok &= good
                    if v is not None:
                        cfg["mcp"][f2] = v
    return cfg, ok


class _Obj:
    """JSON object that keeps member order and duplicates."""

    def __init__(self, items):
        self.items = items

    def last(self, key):
        hit = (False, None)
        for k, v in self.items:
            if k == key:
                hit = (True, v)
        return hit


def parse_effective_safety(payload):   # effective.go:12-39
    if not payload:
        return None
    if isinstance(payload, (bytes, bytearray)):
        try:
            payload = bytes(payload).decode("utf-8")
        except UnicodeDecodeError:
            return None
    try:
        top = json.loads(payload, object_pairs_hook=_Obj,
                         parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
    except ValueError:
        return None
    if not isinstance(top, _Obj):
        return None
    found, raw = top.last("safety")
    if found:
        cfg, ok = _decode_safety(raw)
        if ok:
            return cfg
    found, raw = top.last("data")
    if found and isinstance(raw, _Obj):
        f2, sraw = raw.last("safety")
        if f2:
            cfg, ok = _decode_safety(sraw)
            if ok:
                return cfg
    return None


def match_any(patterns, value) -> bool:   # kernel.go:455-465
    if value == "":
        return False
    return any(glob_ok(p, value) for p in (patterns or []))


# ------------------------------------------------------------------ kernel.evaluate
def pick_label(labels, *keys):   # kernel.go:407-414
    for k in keys:
        if k in labels and trim_space(labels[k]) != "":
            return trim_space(labels[k])
    return ""


def extract_mcp(labels):   # kernel.go:395-405
    if not labels:
        return {f: "" for f in MCP_FIELDS}
    return {"server": pick_label(labels, "mcp.server", "mcp_server", "mcpServer"),
            "tool": pick_label(labels, "mcp.tool", "mcp_tool", "mcpTool"),
            "resource": pick_label(labels, "mcp.resource", "mcp_resource", "mcpResource"),
            "action": to_lower(pick_label(labels, "mcp.action", "mcp_action", "mcpAction"))}


def secrets_present(meta, labels) -> bool:   # kernel.go:381-393
    if labels is not None:
        v = trim_space(labels.get("secrets_present", ""))
        if v != "":
            return v == "true" or v == "1" or equal_fold(v, "yes")
    return any(equal_fold(t, "secrets") for t in meta["risk_tags"])


def policy_meta(job):   # kernel.go:348-379
    meta = job.get("meta")
    out = {"actor_id": "", "actor_type": "", "capability": "", "risk_tags": [], "requires": [], "pack_id": ""}
    if meta is None:
        if job.get("principal_id", "") != "":
            out["actor_id"] = job["principal_id"]
        return out
    out["actor_id"] = meta.get("actor_id", "")
    at = meta.get("actor_type", 0)
    if isinstance(at, str):
        at = {"human": 1, "service": 2}.get(at.lower(), 0)
    out["actor_type"] = {1: "human", 2: "service"}.get(at, "")
    out["capability"] = meta.get("capability", "")
    out["risk_tags"] = list(meta.get("risk_tags") or [])
    out["requires"] = list(meta.get("requires") or [])
    out["pack_id"] = meta.get("pack_id", "")
    if out["actor_id"] == "":
        out["actor_id"] = job.get("principal_id", "")
    return out


def kernel_evaluate(policy, job, gateway=False):   # kernel.go:129-257; gateway=True: policy_bundles.go:1132-1231
    decision, reason = "ALLOW", ""
    topic = trim_space(job.get("topic", ""))
    tenant = trim_space(job.get("tenant", ""))
    meta = job.get("meta")
    if tenant == "" and meta is not None:
        tenant = trim_space(meta.get("tenant_id", ""))
    default_tenant = trim_space(policy.get("default_tenant") or "") if policy is not None else ""
    if tenant == "":
        tenant = default_tenant
    if tenant == "":
        tenant = "default"
    base = {"rule_id": "", "rule_idx": -1, "approval_required": False, "has_snapshot": False,
            "has_constraints": False}
    if topic == "":
        return dict(base, decision="DENY", reason="missing topic")
    if not topic.startswith("job."):
        return dict(base, decision="DENY", reason="unsupported topic")
    labels = job.get("labels") or {}
    inp = {"tenant": tenant, "topic": topic, "labels": labels, "meta": policy_meta(job), "mcp": extract_mcp(labels)}
    inp["secrets_present"] = secrets_present(inp["meta"], labels)
    pd = {"decision": "allow", "reason": "", "rule_id": "", "rule_idx": -1, "approval_required": False,
          "constraints": {}}
    if policy is not None:
        pd = policy_evaluate(policy, inp)
        tenants = policy.get("tenants") or {}
        if tenant in tenants:
            ok, why, _ = mcp_allowed((tenants[tenant] or {}).get("mcp"), inp["mcp"])
            if not ok:
                pd["decision"] = "deny"
                pd["reason"] = why
    has_constraints = not constraints_empty(pd["constraints"])
    d = pd["decision"]
    if d == "deny":
        decision, reason = "DENY", pd["reason"]
    elif d == "require_approval":
        decision, reason = "REQUIRE_HUMAN", pd["reason"]
    elif d == "throttle":
        decision, reason = "THROTTLE", pd["reason"]
    elif d == "allow_with_constraints":
        decision = "ALLOW_WITH_CONSTRAINTS"
    elif d == "allow":
        if has_constraints:
            decision = "ALLOW_WITH_CONSTRAINTS"
    eff = parse_effective_safety(job.get("effective_config"))
    if eff is not None:
        if match_any(eff["denied_topics"], topic):
            decision, reason = "DENY", ("topic %s denied by effective config" % go_quote(topic) if gateway   # :1207
                                        else "topic '%s' denied by effective config" % topic)
        if len(eff["allowed_topics"]) > 0 and not match_any(eff["allowed_topics"], topic):
            decision, reason = "DENY", ("topic %s not allowed by effective config" % go_quote(topic) if gateway   # :1211
                                        else "topic '%s' not allowed by effective config" % topic)
        ok, why, _ = mcp_allowed(eff["mcp"], inp["mcp"])
        if not ok:
            decision, reason = "DENY", why
    approval_required = pd["approval_required"] or decision == "REQUIRE_HUMAN"
    return {"decision": decision, "reason": reason, "rule_id": pd["rule_id"], "rule_idx": pd["rule_idx"],
            "approval_required": approval_required, "has_snapshot": True, "has_constraints": has_constraints}


# ------------------------------------------------------------------ routing
def f32(x) -> float:
    return struct.unpack("<f", struct.pack("<f", x))[0]


def load_score(hb) -> float:   # strategy_least_loaded.go:157-159, float32 left to right
    a = f32(float(hb.get("active_jobs", 0)))
    b = f32(f32(hb.get("cpu_load", 0.0)) / 100.0)
    c = f32(f32(hb.get("gpu_utilization", 0.0)) / 100.0)
    return f32(f32(a + b) + c)


def is_overloaded(hb) -> bool:   # :177-193
    cap = hb.get("max_parallel_jobs", 0)
    if cap > 0:
        if f32(f32(float(hb.get("active_jobs", 0))) / f32(float(cap))) >= f32(0.9):
            return True
    if f32(hb.get("cpu_load", 0.0)) >= 90:
        return True
    if f32(hb.get("gpu_utilization", 0.0)) >= 90:
        return True
    return False


def matches_labels(hb, required) -> bool:   # :161-175
    if not required:
        return True
    labels = hb.get("labels") or {}
    if not labels:
        return False
    return all(labels.get(k, "") == v for k, v in required.items())


_SKIP = {"preferred_worker_id", "preferred_pool", "approval_granted", "secrets_present", "workflow_id", "run_id",
         "step_id", "node_id", "worker_id"}


def filter_placement_labels(labels):   # :195-222
    if not labels:
        return {}
    return {k: v for k, v in labels.items() if k not in _SKIP and not k.startswith("cordum.")}


def pool_satisfies(pool_requires, job_requires) -> bool:   # :241-265
    if not job_requires:
        return True
    if not pool_requires:
        return False
    have = {to_lower(trim_space(r)) for r in pool_requires} - {""}
    for r in job_requires:
        need = to_lower(trim_space(r))
        if need == "":
            continue
        if need not in have:
            return False
    return True


def pick_subject(routing, job, workers):   # :40-136 ; workers: list of heartbeat dicts (slot = index)
    """returns dict(status=..., subject=..., worker_slot=..., tie=bool)."""
    out = {"status": "", "subject": "", "worker_slot": -1, "tie": False}
    topic = job.get("topic", "")
    if topic == "":
        return dict(out, status="missing topic")
    labels = job.get("labels") or {}
    required = filter_placement_labels(labels)
    hint = labels.get("preferred_pool", "")
    topic_pools = list((routing.get("topics") or {}).get(topic) or [])
    if isinstance((routing.get("topics") or {}).get(topic), str):
        topic_pools = [routing["topics"][topic]]
    if hint != "":
        if hint not in topic_pools:
            return dict(out, status="no_pool_mapping:preferred")
        topic_pools = [hint]
    if not topic_pools:
        return dict(out, status="no_pool_mapping:topic")
    job_requires = []
    if job.get("meta") is not None:
        job_requires = list(job["meta"].get("requires") or [])
    pools_cfg = routing.get("pools") or {}
    if not job_requires:
        eligible = topic_pools
    else:
        eligible = [p for p in topic_pools if pool_satisfies((pools_cfg.get(p) or {}).get("requires") or [], job_requires)]
    if not eligible:
        return dict(out, status="no_pool_mapping:requires")
    pool_set = set(eligible)
    registry = {}
    for slot, hb in enumerate(workers):
        registry[hb.get("worker_id", "")] = slot   # map: last wins
    pref = labels.get("preferred_worker_id", "")
    if pref != "" and pref in registry:
        hb = workers[registry[pref]]
        if hb.get("pool", "") in pool_set and matches_labels(hb, required) and not is_overloaded(hb):
            return dict(out, status="ok_preferred", subject="worker.%s.jobs" % pref, worker_slot=registry[pref])
    selected, best, total, over, n_min = None, 0.0, 0, 0, 0
    for wid in sorted(registry.keys(), key=lambda s: s.encode("utf-8", "surrogatepass")):
        slot = registry[wid]
        hb = workers[slot]
        if hb.get("pool", "") not in pool_set:
            continue
        if not matches_labels(hb, required):
            continue
        total += 1
        if is_overloaded(hb):
            over += 1
            continue
        sc = load_score(hb)
        if selected is None or sc < best:
            selected, best, n_min = slot, sc, 1
        elif sc == best:
            n_min += 1
    if selected is None:
        if total > 0 and over == total:
            return dict(out, status="pool_overloaded")
        return dict(out, status="no_workers")
    wid = workers[selected].get("worker_id", "")
    return dict(out, status="ok", subject=("worker.%s.jobs" % wid) if wid != "" else topic, worker_slot=selected,
                tie=n_min > 1)


# ------------------------------------------------------------------ engine glue
def process_job(policy, routing, workers, job):   # engine.go:294-347,393,484-531
    if job.get("approved"):
        pol = {"decision": "ALLOW", "reason": "approval granted", "rule_id": "", "rule_idx": -1,
               "approval_required": False, "has_snapshot": False, "has_constraints": False, "approved_bypass": True}
        sched = "ALLOW"
    else:
        pol = kernel_evaluate(policy, job)
        sched = pol["decision"]
        if pol["approval_required"] and sched in ("ALLOW", "ALLOW_WITH_CONSTRAINTS"):
            sched = "REQUIRE_HUMAN"
    route = None
    if sched in ("ALLOW", "ALLOW_WITH_CONSTRAINTS"):
        route = pick_subject(routing, job, workers)
    return dict(pol, sched_decision=sched, route=route)
