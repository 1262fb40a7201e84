"""Seeded synthetic workloads for BASELINE.json's configs.

  config 2: 10k jobs x 256 rules x 1k workers            make_config("c2")
  config 3: 1M jobs x 4096 rules x 64k workers           make_config("c3")
  config 5: demo-guardrails replay, 100k jobs            make_c5(...)

Everything is a pure function of the seed.  Outputs are the boundary documents /
wire tables (policy dict, routing dict, WorkerTable, EnvelopeBatch); no policy logic
lives here.  Vectorized with numpy so the 1M-job config builds in a few seconds.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

from . import wire

VERBS = ["read", "write", "delete", "deploy", "scan", "build", "query", "sync"]
LABEL_KEYS = ["region", "gpu", "zone", "arch", "tier", "disk", "net", "os"]


@dataclass
class Spec:
    name: str
    n_jobs: int
    n_rules: int
    n_workers: int
    seed: int
    n_tenants: int
    n_packs: int
    n_caps: int
    n_risk: int
    n_requires: int
    n_pools: int
    n_label_vals: int   # label pairs = len(LABEL_KEYS) x n_label_vals
    n_actors: int = 4096
    n_effcfg: int = 16
    tie_fraction: float = 0.0
    edge_fraction: float = 0.002


SPECS = {
    "tiny": Spec("tiny", 2000, 64, 200, 1, 8, 8, 16, 16, 16, 8, 2, n_actors=64, n_effcfg=4),
    "c2": Spec("c2", 10_000, 256, 1_000, 2, 16, 16, 64, 32, 32, 16, 4),
    "c3": Spec("c3", 1_000_000, 4096, 65_536, 3, 256, 256, 512, 64, 64, 256, 8, n_effcfg=64),
}


@dataclass
class Config:
    spec: Spec
    policy: dict
    routing: dict
    workers: wire.WorkerTable
    jobs: wire.EnvelopeBatch
    info: dict = field(default_factory=dict)


class _Dict:
    """A list of strings laid into a shared arena; .spans[i] = (off,len) of string i."""

    def __init__(self, arena: bytearray, strings):
        self.strings = list(strings)
        self.spans = np.zeros(len(self.strings), dtype=wire.STR_DTYPE)
        for i, s in enumerate(self.strings):
            b = s.encode("utf-8")
            self.spans[i] = (len(arena), len(b))
            arena += b

    def __len__(self):
        return len(self.strings)


def _zipf_choice(rng, n, size, a=1.1):
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), a)
    return rng.choice(n, size=size, p=w / w.sum())


def _mk_policy(spec: Spec, rng, D) -> dict:
    tenants, topics, packs, caps, risk, reqs, actors, label_pairs = (
        D["tenants"], D["topics"], D["packs"], D["caps"], D["risk"], D["reqs"], D["actors"], D["policy_labels"])
    rules = []
    decisions = ["allow"] * 50 + ["deny"] * 20 + ["require_approval"] * 15 + ["allow_with_constraints"] * 10 + ["throttle"] * 5

    def pick(lst, k):
        idx = rng.choice(len(lst), size=min(k, len(lst)), replace=False)
        return [lst[i] for i in idx]

    def topic_pattern():
        r = rng.random()
        if r < 0.55:
            return topics[rng.integers(len(topics))]
        if r < 0.88:
            return "job.%s.*" % packs[rng.integers(len(packs))]
        if r < 0.90:
            return "job.*.%s" % VERBS[rng.integers(len(VERBS))]
        if r < 0.94:
            return "job.%s.[a-m]*" % packs[rng.integers(len(packs))]
        if r < 0.97:
            return "job.%s.?????" % packs[rng.integers(len(packs))]
        if r < 0.985:
            return " job.%s.* " % packs[rng.integers(len(packs))]   # padded: trimmed by matchTopic
        return "job.[invalid"                                        # malformed: never matches

    def constraints():
        c = {}
        if rng.random() < 0.7:
            c["budgets"] = {"max_runtime_ms": int(rng.integers(1, 600_000)), "max_retries": int(rng.integers(0, 5)),
                            "max_concurrent_jobs": int(rng.integers(0, 16))}
        if rng.random() < 0.5:
            c["sandbox"] = {"isolated": bool(rng.random() < 0.5), "network_allowlist": ["api.internal"],
                            "fs_read_only": ["/etc"]}
        if rng.random() < 0.3:
            c["toolchain"] = {"allowed_tools": ["git", "make"], "allowed_commands": ["build"]}
        if rng.random() < 0.3:
            c["diff"] = {"max_files": int(rng.integers(1, 50)), "max_lines": int(rng.integers(10, 2000)),
                         "deny_path_globs": ["secrets/**"]}
        if not c:
            c["redaction_level"] = "strict"
        return c

    for i in range(spec.n_rules - 1):
        m = {}
        # at least one selective predicate so the first-match depth is spread over the table
        sel = rng.random()
        if sel < 0.12 or sel >= 0.30:
            m["tenants"] = pick(tenants, int(rng.integers(1, 4)))
        if sel >= 0.12:
            m["topics"] = [topic_pattern() for _ in range(int(rng.integers(1, 3)))]
        extra = int(rng.integers(0, 4))
        opts = ["capabilities", "risk_tags", "requires", "pack_ids", "actor_ids", "actor_types", "secrets_present",
                "labels", "mcp"]
        w = np.array([0.22, 0.18, 0.12, 0.15, 0.04, 0.08, 0.08, 0.08, 0.05])
        for name in rng.choice(opts, size=extra, replace=False, p=w / w.sum()):
            if name == "capabilities":
                m[name] = pick(caps, int(rng.integers(1, 4)))
                if rng.random() < 0.1:
                    m[name][0] = " " + m[name][0].upper() + " "   # case/space-insensitive match
            elif name == "risk_tags":
                m[name] = pick(risk, int(rng.integers(1, 3)))
            elif name == "requires":
                m[name] = pick(reqs, int(rng.integers(1, 3)))
            elif name == "pack_ids":
                m[name] = pick(packs, int(rng.integers(1, 3)))
            elif name == "actor_ids":
                m[name] = pick(actors, int(rng.integers(1, 3)))
            elif name == "actor_types":
                m[name] = [["human", "service", "HUMAN"][rng.integers(3)]]
            elif name == "secrets_present":
                m[name] = bool(rng.random() < 0.5)
            elif name == "labels":
                k, v = label_pairs[rng.integers(len(label_pairs))]
                m[name] = {k: v}
            elif name == "mcp":
                mc = {}
                if rng.random() < 0.5:
                    mc["allow_servers"] = pick(D["mcp_servers"], 2)
                if rng.random() < 0.5:
                    mc["deny_tools"] = pick(D["mcp_tools"], 1)
                if rng.random() < 0.3:
                    mc["allow_actions"] = ["read", "list"]
                m[name] = mc
        rule = {"id": "rule-%05d" % i, "match": m, "decision": decisions[rng.integers(len(decisions))],
                "reason": "reason for rule %d" % i}
        if rule["decision"] == "allow_with_constraints" or rng.random() < 0.18:
            rule["constraints"] = constraints()
        if rule["decision"] == "deny" and rng.random() < 0.3:
            rule["remediations"] = [{"id": "alt-%d" % i, "title": "Use the safe path", "replacement_topic": topics[0]}]
        rules.append(rule)
    half = [tenants[i] for i in range(0, len(tenants), 2)]
    rules.append({"id": "catch-all", "match": {"tenants": half}, "decision": "allow", "reason": "catch-all"})
    tenant_cfg = {}
    for t in tenants[: max(2, len(tenants) // 4)]:
        mc = {}
        if rng.random() < 0.5:
            mc["deny_servers"] = pick(D["mcp_servers"], 1)
        if rng.random() < 0.3:
            mc["allow_tools"] = pick(D["mcp_tools"], 3)
        tenant_cfg[t] = {"allow_topics": ["job.*"], "deny_topics": [], "mcp": mc}
    return {"version": "synthetic-%s" % spec.name, "default_tenant": tenants[0], "rules": rules, "tenants": tenant_cfg}


def _mk_routing(spec: Spec, rng, D) -> dict:
    pools, topics, reqs = D["pools"], D["topics"], D["reqs"]
    pool_cfg = {}
    for p in pools:
        k = int(rng.integers(0, 4))
        pool_cfg[p] = {"requires": [reqs[i] for i in rng.choice(len(reqs), size=k, replace=False)]}
    # a few broad pools so that jobs with requires can usually be placed
    for p in pools[: max(1, len(pools) // 8)]:
        pool_cfg[p] = {"requires": list(reqs[: max(4, len(reqs) // 2)])}
    tmap = {}
    for i, t in enumerate(topics):
        if rng.random() < 0.03:
            continue   # unmapped topic -> ErrNoPoolMapping
        k = int(rng.integers(1, 4))
        lst = [pools[j] for j in rng.choice(len(pools), size=k, replace=False)]
        if rng.random() < 0.5:
            lst.append(pools[int(rng.integers(0, max(1, len(pools) // 8)))])
        tmap[t] = lst
    return {"topics": tmap, "pools": pool_cfg}


def _csr(n, job_idx, *cols):
    """Group entries (job_idx, cols...) into CSR over n jobs (stable)."""
    order = np.argsort(job_idx, kind="stable")
    counts = np.bincount(job_idx, minlength=n).astype(np.uint32)
    off = np.zeros(n + 1, np.uint32)
    np.cumsum(counts, out=off[1:])
    out = []
    for c in cols:
        c = c[order]
        out.append(c if len(c) else np.zeros(1, dtype=wire.STR_DTYPE))
    return off, out


def make_config(name_or_spec, n_jobs: int | None = None) -> Config:
    spec = SPECS[name_or_spec] if isinstance(name_or_spec, str) else name_or_spec
    if n_jobs is not None:
        spec = Spec(**{**spec.__dict__, "n_jobs": n_jobs})
    rng = np.random.default_rng(spec.seed)
    arena = bytearray(b"\0")   # offset 0 reserved so (0,0) is the empty string

    packs = ["pack%03d" % i for i in range(spec.n_packs)]
    topics = ["job.%s.%s" % (p, v) for p in packs for v in VERBS]
    pol_pairs = [("env", v) for v in ("prod", "staging", "dev")] + [("team", "t%d" % i) for i in range(5)]
    place_pairs = [(k, "%s-%d" % (k[0], i)) for k in LABEL_KEYS for i in range(spec.n_label_vals)]
    S = dict(
        tenants=["tenant-%03d" % i for i in range(spec.n_tenants)],
        packs=packs, topics=topics,
        caps=["cap.%03d" % i for i in range(spec.n_caps)],
        risk=["risk%02d" % i for i in range(spec.n_risk - 2)] + ["write", "secrets"],
        reqs=["req%02d" % i for i in range(spec.n_requires)],
        pools=["pool-%03d" % i for i in range(spec.n_pools)],
        actors=["user-%05d" % i for i in range(spec.n_actors)],
        mcp_servers=["mcp%02d.example.com" % i for i in range(12)],
        mcp_tools=["tool%02d" % i for i in range(12)],
        mcp_resources=["res://bucket/%02d" % i for i in range(8)],
        mcp_actions=["read", "LIST", "write", "Delete"],
        policy_labels=pol_pairs,
    )
    policy = _mk_policy(spec, rng, S)
    routing = _mk_routing(spec, rng, S)

    # ---------------------------------------------------------------- workers
    W = spec.n_workers
    wid = _Dict(arena, ["w-%06d" % i for i in range(W)])
    pool_d = _Dict(arena, S["pools"] + ["pool-unrouted"])
    wpool = _zipf_choice(rng, spec.n_pools, W, a=0.6)
    wpool[rng.random(W) < 0.01] = spec.n_pools           # a few workers in a pool no topic maps to
    active = rng.integers(0, 9, W).astype(np.int32)
    maxp = rng.choice(np.array([0, 4, 8, 16], np.int32), W)
    cpu = (rng.random(W) * 100.0).astype(np.float32)
    gpu = (rng.random(W) * 100.0).astype(np.float32)
    if spec.tie_fraction > 0:
        t = rng.random(W) < spec.tie_fraction
        active[t], cpu[t], gpu[t], maxp[t] = 1, 25.0, 50.0, 0
    place_d = _Dict(arena, [k for k, _ in place_pairs] + [v for _, v in place_pairs])
    npairs = len(place_pairs)
    polpair_d = _Dict(arena, [k for k, _ in pol_pairs] + [v for _, v in pol_pairs])
    # worker labels: up to 3 distinct placement keys + sometimes env=...
    wl_job, wl_k, wl_v = [], [], []
    nkeys = len(LABEL_KEYS)
    wk = rng.random((W, nkeys)).argsort(axis=1)[:, :3]             # 3 distinct keys per worker
    wn = rng.integers(0, 4, W)                                     # 0..3 labels
    for c in range(3):
        sel = np.nonzero(wn > c)[0]
        key = wk[sel, c]
        val = rng.integers(0, spec.n_label_vals, len(sel))
        pair = key * spec.n_label_vals + val
        wl_job.append(sel)
        wl_k.append(place_d.spans[pair])
        wl_v.append(place_d.spans[npairs + pair])
    sel = np.nonzero(rng.random(W) < 0.3)[0]
    envv = rng.integers(0, 3, len(sel))
    wl_job.append(sel)
    wl_k.append(polpair_d.spans[envv])
    wl_v.append(polpair_d.spans[len(pol_pairs) + envv])
    wl_off, (wl_keys, wl_vals) = _csr(W, np.concatenate(wl_job), np.concatenate(wl_k), np.concatenate(wl_v))

    # ---------------------------------------------------------------- jobs
    J = spec.n_jobs
    ten_d = _Dict(arena, S["tenants"] + [t.upper() for t in S["tenants"][:4]] + ["  " + S["tenants"][1] + " ", "unknown-tenant"])
    top_d = _Dict(arena, topics + ["", "sys.destroy", "  " + topics[0] + "  ", "job.unlisted.topic", "job."])
    cap_d = _Dict(arena, S["caps"] + ["cap.unreferenced", " CAP.001 "])
    pack_d = _Dict(arena, packs + ["pack-unknown"])
    actor_d = _Dict(arena, S["actors"])
    risk_d = _Dict(arena, S["risk"] + ["unreferenced-tag", "WRITE", "Secrets"])
    req_d = _Dict(arena, S["reqs"] + ["REQ00", " req01 ", "req-unknown", " "])
    misc = _Dict(arena, ["workflow_id", "run_id", "step_id", "preferred_pool", "preferred_worker_id", "secrets_present",
                         "mcp.server", "mcp_tool", "mcpResource", "mcp.action", "true", "no", "1", "yes",
                         "wf-1", "run-1", "step-1", "pool-nonexistent", "w-nonexistent", "cordum.trace", "t-1"])
    mi = {s: i for i, s in enumerate(misc.strings)}
    mcp_srv = _Dict(arena, S["mcp_servers"] + ["other.example.com"])
    mcp_tool = _Dict(arena, S["mcp_tools"] + ["tool-other"])
    mcp_res = _Dict(arena, S["mcp_resources"])
    mcp_act = _Dict(arena, S["mcp_actions"])
    effcfgs = []
    for i in range(spec.n_effcfg):
        s = {}
        r = i % 4
        if r in (0, 2):
            s["denied_topics"] = ["job.%s.*" % packs[(7 * i) % len(packs)], topics[(13 * i) % len(topics)]]
        if r in (1, 2):
            s["allowed_topics"] = ["job.%s.*" % packs[(3 * i + k) % len(packs)] for k in range(max(2, len(packs) // 2))]
        if r == 3:
            s["mcp"] = {"deny_servers": [S["mcp_servers"][i % 12]], "allow_actions": ["read", "list"]}
        doc = {"safety": s} if i % 3 else {"data": {"safety": s}}
        effcfgs.append(json.dumps(doc))
    effcfgs.append('{"safety": {"denied_topics": "not-a-list"}}')   # type error -> ignored
    effcfgs.append("{not json")                                     # unparsable -> ignored
    eff_d = _Dict(arena, effcfgs)

    jt = _zipf_choice(rng, spec.n_tenants, J)
    jtopic = _zipf_choice(rng, len(topics), J)
    # spread popular topics over packs rather than clustering on pack000
    perm = rng.permutation(len(topics))
    jtopic = perm[jtopic]
    tenant_ix = jt.copy()
    u = rng.random(J)
    tenant_ix[u < 0.01] = spec.n_tenants + rng.integers(0, 4, int((u < 0.01).sum()))   # upper-case variants
    tenant_ix[(u >= 0.01) & (u < 0.012)] = spec.n_tenants + 4                              # padded
    tenant_ix[(u >= 0.012) & (u < 0.02)] = spec.n_tenants + 5                              # unknown
    topic_ix = jtopic.copy()
    e = rng.random(J)
    ef = spec.edge_fraction
    nt = len(topics)
    topic_ix[e < ef * 0.2] = nt + 0
    topic_ix[(e >= ef * 0.2) & (e < ef * 0.4)] = nt + 1
    topic_ix[(e >= ef * 0.4) & (e < ef * 0.6)] = nt + 2
    topic_ix[(e >= ef * 0.6) & (e < ef * 0.8)] = nt + 3
    topic_ix[(e >= ef * 0.8) & (e < ef)] = nt + 4
    has_meta = (rng.random(J) >= 0.02).astype(np.uint8)
    cap_ix = rng.integers(0, spec.n_caps, J)
    c = rng.random(J)
    cap_ix[c < 0.05] = spec.n_caps          # unreferenced capability
    cap_ix[(c >= 0.05) & (c < 0.06)] = spec.n_caps + 1
    pack_ix = (jtopic // len(VERBS)).astype(np.int64)
    p = rng.random(J)
    pack_ix[p < 0.08] = rng.integers(0, spec.n_packs, int((p < 0.08).sum()))
    pack_ix[(p >= 0.08) & (p < 0.10)] = spec.n_packs
    actor_ix = rng.integers(0, spec.n_actors, J)
    actor_type = rng.integers(0, 3, J).astype(np.uint8)
    principal = actor_d.spans[rng.integers(0, spec.n_actors, J)].copy()
    actor_sp = actor_d.spans[actor_ix].copy()
    noactor = rng.random(J) < 0.1
    actor_sp[noactor] = (0, 0)              # falls back to principal (kernel.go:364)
    empty = np.zeros(J, dtype=wire.STR_DTYPE)
    hm = has_meta.astype(bool)

    # risk tags 0..3
    nr = rng.choice(4, J, p=[0.35, 0.35, 0.2, 0.1])
    nr[~hm] = 0
    r_job = np.repeat(np.arange(J), nr)
    r_val = rng.integers(0, len(risk_d), len(r_job))
    risk_off, (risk_sp,) = _csr(J, r_job, risk_d.spans[r_val])
    # requires 0..2
    nq = rng.choice(3, J, p=[0.6, 0.3, 0.1])
    nq[~hm] = 0
    q_job = np.repeat(np.arange(J), nq)
    q_val = rng.integers(0, max(4, spec.n_requires // 2), len(q_job))
    odd = rng.random(len(q_job)) < 0.03
    q_val[odd] = spec.n_requires + rng.integers(0, 4, int(odd.sum()))
    req_off, (req_sp,) = _csr(J, q_job, req_d.spans[q_val])

    # labels
    lj, lk, lv = [], [], []

    def add(sel, kspans, vspans):
        if len(sel):
            lj.append(sel)
            lk.append(kspans)
            lv.append(vspans)

    def rep(span, n):
        a = np.zeros(n, dtype=wire.STR_DTYPE)
        a[:] = span
        return a

    sel = np.nonzero(rng.random(J) < 0.5)[0]
    for k, v in (("workflow_id", "wf-1"), ("run_id", "run-1"), ("step_id", "step-1")):
        add(sel, rep(misc.spans[mi[k]], len(sel)), rep(misc.spans[mi[v]], len(sel)))
    sel = np.nonzero(rng.random(J) < 0.05)[0]
    add(sel, rep(misc.spans[mi["cordum.trace"]], len(sel)), rep(misc.spans[mi["t-1"]], len(sel)))
    # placement labels: 30 % of jobs, 1-2 distinct keys
    pl = rng.random(J)
    jk = rng.random((J, nkeys)).argsort(axis=1)[:, :2] if J <= 200_000 else np.stack(
        [rng.integers(0, nkeys, J), rng.integers(0, nkeys, J)], axis=1)
    if J > 200_000:
        same = jk[:, 0] == jk[:, 1]
        jk[same, 1] = (jk[same, 1] + 1) % nkeys
    for c_, thr in ((0, 0.30), (1, 0.12)):
        sel = np.nonzero(pl < thr)[0]
        key = jk[sel, c_]
        val = rng.integers(0, spec.n_label_vals, len(sel))
        pair = key * spec.n_label_vals + val
        add(sel, place_d.spans[pair], place_d.spans[npairs + pair])
    # policy labels env=/team= : 10 % (they also constrain placement — reference quirk)
    sel = np.nonzero(rng.random(J) < 0.10)[0]
    pv = rng.integers(0, len(pol_pairs), len(sel))
    keep_env = pv < 3
    add(sel[keep_env], polpair_d.spans[pv[keep_env]], polpair_d.spans[len(pol_pairs) + pv[keep_env]])
    add(sel[~keep_env], polpair_d.spans[pv[~keep_env]], polpair_d.spans[len(pol_pairs) + pv[~keep_env]])
    # secrets_present label 2 %
    sel = np.nonzero(rng.random(J) < 0.02)[0]
    sv_ = np.array([mi["true"], mi["no"], mi["1"], mi["yes"]])[rng.integers(0, 4, len(sel))]
    add(sel, rep(misc.spans[mi["secrets_present"]], len(sel)), misc.spans[sv_])
    # MCP 5 %
    m = rng.random(J)
    sel = np.nonzero(m < 0.05)[0]
    add(sel, rep(misc.spans[mi["mcp.server"]], len(sel)), mcp_srv.spans[rng.integers(0, len(mcp_srv), len(sel))])
    sel2 = sel[rng.random(len(sel)) < 0.7]
    add(sel2, rep(misc.spans[mi["mcp_tool"]], len(sel2)), mcp_tool.spans[rng.integers(0, len(mcp_tool), len(sel2))])
    sel3 = sel[rng.random(len(sel)) < 0.3]
    add(sel3, rep(misc.spans[mi["mcpResource"]], len(sel3)), mcp_res.spans[rng.integers(0, len(mcp_res), len(sel3))])
    sel4 = sel[rng.random(len(sel)) < 0.5]
    add(sel4, rep(misc.spans[mi["mcp.action"]], len(sel4)), mcp_act.spans[rng.integers(0, len(mcp_act), len(sel4))])
    # preferred_pool 1 %, preferred_worker_id 1 %
    sel = np.nonzero(rng.random(J) < 0.01)[0]
    pp = pool_d.spans[rng.integers(0, spec.n_pools, len(sel))].copy()
    bad = rng.random(len(sel)) < 0.1
    pp[bad] = misc.spans[mi["pool-nonexistent"]]
    add(sel, rep(misc.spans[mi["preferred_pool"]], len(sel)), pp)
    sel = np.nonzero(rng.random(J) < 0.01)[0]
    pw = wid.spans[rng.integers(0, W, len(sel))].copy()
    bad = rng.random(len(sel)) < 0.1
    pw[bad] = misc.spans[mi["w-nonexistent"]]
    add(sel, rep(misc.spans[mi["preferred_worker_id"]], len(sel)), pw)
    lab_off, (lab_k, lab_v) = _csr(J, np.concatenate(lj), np.concatenate(lk), np.concatenate(lv))

    # effective config 2 %
    eff = empty.copy()
    sel = np.nonzero(rng.random(J) < 0.02)[0]
    eff[sel] = eff_d.spans[rng.integers(0, len(eff_d), len(sel))]

    def meta_col(spans):
        out = spans.copy()
        out[~hm] = (0, 0)
        return out

    cols = dict(
        topic=top_d.spans[topic_ix].copy(), tenant=ten_d.spans[tenant_ix].copy(), principal_id=principal,
        effective_config=eff, has_meta=has_meta, meta_tenant_id=empty.copy(), actor_id=meta_col(actor_sp),
        actor_type=np.where(hm, actor_type, 0).astype(np.uint8), capability=meta_col(cap_d.spans[cap_ix]),
        pack_id=meta_col(pack_d.spans[pack_ix]), risk_off=risk_off, risk_tags=risk_sp, requires_off=req_off,
        requires_=req_sp, label_off=lab_off, label_keys=lab_k, label_vals=lab_v, approved=np.zeros(J, np.uint8))
    # a few jobs leave Tenant empty and rely on meta.tenant_id / default_tenant (kernel.go:136-169)
    sel = np.nonzero(rng.random(J) < 0.01)[0]
    cols["tenant"][sel] = (0, 0)
    half = sel[: len(sel) // 2]
    cols["meta_tenant_id"][half] = ten_d.spans[jt[half]]
    cols["meta_tenant_id"][~hm] = (0, 0)

    arena_np = np.frombuffer(bytes(arena), dtype=np.uint8).copy()
    workers = wire.WorkerTable(W, arena_np, dict(
        worker_id=wid.spans, pool=pool_d.spans[wpool].copy(), active_jobs=active, max_parallel_jobs=maxp, cpu_load=cpu,
        gpu_utilization=gpu, label_off=wl_off, label_keys=wl_keys, label_vals=wl_vals))
    jobs = wire.EnvelopeBatch(J, arena_np, cols)
    return Config(spec, policy, routing, workers, jobs, info={"n_topics": len(topics), "n_effcfg": len(effcfgs)})


def make_c5(n_jobs: int = 100_000, seed: int = 5, policy=None, routing=None):
    """demo-guardrails replay: 70 % write+[write,prod] (REQUIRE_APPROVAL),
    10 % write untagged (default ALLOW), 10 % dangerous (DENY), 10 % safe (ALLOW); `approved`
    marks the replay of the approved 70 % (engine.go:484-522)."""
    rng = np.random.default_rng(seed)
    arena = bytearray(b"\0")
    d = _Dict(arena, ["job.demo-guardrails.write", "job.demo-guardrails.dangerous", "job.demo-guardrails.safe",
                      "default", "demo-guardrails", "demo-guardrails.write", "demo-guardrails.dangerous",
                      "demo-guardrails.safe", "write", "prod", "approval_granted", "true", "workflow_id", "wf-demo"])
    ix = {s: i for i, s in enumerate(d.strings)}
    J = n_jobs
    kind = rng.choice(4, J, p=[0.7, 0.1, 0.1, 0.1])   # 0 write tagged, 1 write untagged, 2 dangerous, 3 safe
    topic = d.spans[np.array([0, 0, 1, 2])[kind]].copy()
    cap = d.spans[np.array([5, 5, 6, 7])[kind]].copy()
    nr = np.where(kind == 0, 2, 0)
    r_job = np.repeat(np.arange(J), nr)
    r_val = np.tile(np.array([ix["write"], ix["prod"]]), int((kind == 0).sum()))
    risk_off, (risk_sp,) = _csr(J, r_job, d.spans[r_val])
    lab_off, (lab_k, lab_v) = _csr(J, np.arange(J), np.repeat(d.spans[ix["workflow_id"]:ix["workflow_id"] + 1], J),
                                   np.repeat(d.spans[ix["wf-demo"]:ix["wf-demo"] + 1], J))
    empty = np.zeros(J, dtype=wire.STR_DTYPE)
    one = np.zeros(1, dtype=wire.STR_DTYPE)

    def rep(i):
        a = np.zeros(J, dtype=wire.STR_DTYPE)
        a[:] = d.spans[i]
        return a

    cols = dict(topic=topic, tenant=rep(ix["default"]), principal_id=empty.copy(), effective_config=empty.copy(),
                has_meta=np.ones(J, np.uint8), meta_tenant_id=empty.copy(), actor_id=empty.copy(),
                actor_type=np.zeros(J, np.uint8), capability=cap, pack_id=rep(ix["demo-guardrails"]),
                risk_off=risk_off, risk_tags=risk_sp, requires_off=np.zeros(J + 1, np.uint32), requires_=one,
                label_off=lab_off, label_keys=lab_k, label_vals=lab_v, approved=np.zeros(J, np.uint8))
    arena_np = np.frombuffer(bytes(arena), dtype=np.uint8).copy()
    jobs = wire.EnvelopeBatch(J, arena_np, cols)
    workers = wire.WorkerTable.from_workers(
        [{"worker_id": "demo-worker-%d" % i, "pool": "demo-guardrails", "active_jobs": int(i % 3),
          "max_parallel_jobs": 8, "cpu_load": float(5 * i), "gpu_utilization": 0.0} for i in range(4)])
    return jobs, workers, kind
