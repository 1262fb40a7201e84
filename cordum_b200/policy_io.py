"""Host-side loading of the two documents the engine compiles into device tables.

Mirrors (setup code, not the hot path):
  config.ParseSafetyPolicy            core/infra/config/safety_policy.go:168-184
  safetykernel.mergePolicies          core/controlplane/safetykernel/kernel.go:694-778
  config.ParsePoolsConfig             core/infra/config/pools.go:28-131
  buildRouting                        cmd/cordum-scheduler/config_overlay.go:155-174
The result is plain dicts keyed by the reference's yaml tag names; `json.dumps` of
them is exactly what cordum_policy_load / cordum_routing_load take.
"""
from __future__ import annotations

import copy
import json

import yaml

VALID_DECISIONS = {"allow", "deny", "require_approval", "allow_with_constraints", "throttle"}
_MCP_KEYS = ("allow_servers", "deny_servers", "allow_tools", "deny_tools", "allow_resources", "deny_resources",
             "allow_actions", "deny_actions")


def parse_safety_policy(text) -> dict | None:
    """ParseSafetyPolicy: empty -> None (allow-all); decision enum checked like the embedded
    JSON schema does (config/schema/safety_policy.schema.json:25-28)."""
    if text is None:
        return None
    if isinstance(text, (bytes, bytearray)):
        text = bytes(text).decode("utf-8")
    if len(text) == 0:
        return None
    doc = yaml.safe_load(text)
    if doc is None:
        doc = {}
    if not isinstance(doc, dict):
        raise ValueError("safety policy: expected a mapping")
    for rule in doc.get("rules") or []:
        dec = rule.get("decision")
        if dec is not None and dec not in VALID_DECISIONS:
            raise ValueError("safety policy: invalid decision %r" % (dec,))
    if doc.get("tenants") is None:
        doc["tenants"] = {}
    return doc


def _clone_tenant(tp: dict | None) -> dict:
    tp = tp or {}
    return {
        "allow_topics": list(tp.get("allow_topics") or []),
        "deny_topics": list(tp.get("deny_topics") or []),
        "allowed_repo_hosts": list(tp.get("allowed_repo_hosts") or []),
        "denied_repo_hosts": list(tp.get("denied_repo_hosts") or []),
        "max_concurrent_jobs": int(tp.get("max_concurrent_jobs") or 0),
        "mcp": copy.deepcopy(tp.get("mcp") or {}),
    }


def _clone_policy(p: dict | None) -> dict | None:
    if p is None:
        return None
    return {
        "version": p.get("version") or "",
        "default_tenant": p.get("default_tenant") or "",
        "rules": copy.deepcopy(list(p.get("rules") or [])),
        "tenants": {k: _clone_tenant(v) for k, v in (p.get("tenants") or {}).items()},
    }


def merge_policies(base: dict | None, extra: dict | None) -> dict | None:
    """mergePolicies (kernel.go:694-711): fragment rules are appended AFTER base rules;
    tenants merged list-wise (:731-754); MCP lists concatenated (:767-778)."""
    if base is None:
        return _clone_policy(extra)
    if extra is None:
        return _clone_policy(base)
    out = _clone_policy(base)
    if out["version"] == "":
        out["version"] = extra.get("version") or ""
    if out["default_tenant"] == "":
        out["default_tenant"] = extra.get("default_tenant") or ""
    out["rules"] = out["rules"] + copy.deepcopy(list(extra.get("rules") or []))
    tenants = out["tenants"]
    for name, add in (extra.get("tenants") or {}).items():
        if name not in tenants:
            tenants[name] = _clone_tenant(add)
            continue
        cur = tenants[name]
        add = add or {}
        for k in ("allow_topics", "deny_topics", "allowed_repo_hosts", "denied_repo_hosts"):
            cur[k] = cur[k] + list(add.get(k) or [])
        amax = int(add.get("max_concurrent_jobs") or 0)
        if amax > 0 and (cur["max_concurrent_jobs"] == 0 or amax < cur["max_concurrent_jobs"]):
            cur["max_concurrent_jobs"] = amax
        am = add.get("mcp") or {}
        cur["mcp"] = {k: list((cur["mcp"] or {}).get(k) or []) + list(am.get(k) or []) for k in _MCP_KEYS}
    return out


def parse_pools_config(text) -> dict:
    """ParsePoolsConfig + buildRouting -> {"topics": {t: [pools]}, "pools": {p: {"requires": [...]}}}."""
    if isinstance(text, (bytes, bytearray)):
        text = bytes(text).decode("utf-8")
    raw = yaml.safe_load(text) if isinstance(text, str) else text
    raw = raw or {}
    topics = {}
    for topic, value in (raw.get("topics") or {}).items():
        if topic == "":
            raise ValueError("invalid topic mapping: empty topic")
        if isinstance(value, str):
            if value == "":
                raise ValueError("invalid topic mapping: %r -> empty pool" % topic)
            topics[topic] = [value]
        elif isinstance(value, list):
            pools = []
            for item in value:
                if not isinstance(item, str) or item == "":
                    raise ValueError("invalid pool list for topic %r" % topic)
                pools.append(item)
            if not pools:
                raise ValueError("invalid topic mapping: %r -> empty pools" % topic)
            topics[topic] = pools
        else:
            raise ValueError("invalid topic mapping for %r" % topic)
    if not topics:
        raise ValueError("pool config has no topics")
    pools = {}
    for name, cfg in (raw.get("pools") or {}).items():
        pools[name] = {"requires": list((cfg or {}).get("requires") or [])}
    return {"topics": topics, "pools": pools}


def json_merge_patch(target, patch):
    """RFC 7386, the strategy pack overlays use for pools/timeouts (pack.yaml `json_merge_patch`)."""
    if not isinstance(patch, dict):
        return copy.deepcopy(patch)
    out = copy.deepcopy(target) if isinstance(target, dict) else {}
    for k, v in patch.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = json_merge_patch(out.get(k), v)
    return out


def to_json(doc) -> bytes:
    return b"" if doc is None else json.dumps(doc, separators=(",", ":")).encode("utf-8")
