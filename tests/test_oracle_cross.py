"""Randomized cross-check of the two independent restatements (C++ oracle vs Python
oracle) — the substitute for the un-runnable Go reference on inputs the reference's
own tests do not pin (SURVEY.md §8c last row).  CPU only."""
import fnmatch
import os
import random
import sys

import numpy as np
import pytest

import kats
import oracle_lib
from cordum_b200 import synth, wire

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import py_oracle  # noqa: E402

STATUS_CODE = {v: k for k, v in kats.ROUTE_NAMES.items()}


def py_records(cfg, jobs, mode=wire.MODE_POLICY_AND_ROUTE):
    workers = cfg.workers.to_workers()
    out = []
    for job in jobs:
        r = py_oracle.process_job(cfg.policy, cfg.routing, workers, job)
        rt = r["route"]
        out.append((wire.DEC_NAMES.index(r["decision"]), wire.DEC_NAMES.index(r["sched_decision"]), r["rule_idx"],
                    bool(r["approval_required"]), bool(r["has_snapshot"]), bool(r["has_constraints"]),
                    STATUS_CODE[rt["status"]] if rt else 0, rt["worker_slot"] if rt else -1,
                    bool(rt["tie"]) if rt else False))
    return out


def cpp_records(cfg, env):
    o = oracle_lib.Oracle(cfg.policy, cfg.routing, cfg.workers)
    rec = o.eval(env, wire.MODE_POLICY_AND_ROUTE, threads=4)
    o.close()
    f = rec["flags"]
    return [(int(r["decision"]), int(r["sched_decision"]), int(r["rule_idx"]), bool(fl & wire.F_APPROVAL_REQUIRED),
             bool(fl & wire.F_HAS_SNAPSHOT), bool(fl & wire.F_CONSTRAINTS), int(r["route_status"]),
             int(r["worker_slot"]), bool(fl & wire.F_TIE)) for r, fl in zip(rec, f)], rec


@pytest.mark.parametrize("name,n", [("tiny", 2000), ("c2", 1200)])
def test_restatements_agree_on_synthetic_config(name, n):
    cfg = synth.make_config(name, n)
    a, rec = cpp_records(cfg, cfg.jobs)
    b = py_records(cfg, cfg.jobs.to_jobs())
    bad = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert not bad, bad[:5]
    # the sample must exercise the interesting branches, or agreement means little
    assert len(set(rec["decision"].tolist())) >= 4
    assert len(set(rec["route_status"].tolist())) >= 4
    assert (rec["reason_code"] >= wire.REASON_TENANT_MCP).any()


def test_restatements_agree_with_injected_ties():
    spec = synth.Spec(**{**synth.SPECS["tiny"].__dict__, "tie_fraction": 0.5, "seed": 11})
    cfg = synth.make_config(spec)
    a, rec = cpp_records(cfg, cfg.jobs)
    b = py_records(cfg, cfg.jobs.to_jobs())
    assert a == b
    assert (rec["flags"] & wire.F_TIE).astype(bool).sum() > 20   # ties really occur


def test_approved_replay_agrees():
    cfg = synth.make_config("tiny", 600)
    mask = (np.arange(600) % 3 == 0).astype(np.uint8)
    env = cfg.jobs.with_approved(mask)
    a, rec = cpp_records(cfg, env)
    b = py_records(cfg, env.to_jobs())
    assert a == b
    assert (rec["flags"] & wire.F_APPROVED_BYPASS).astype(bool).sum() == int(mask.sum())


def test_path_match_fuzz_cpp_vs_python():
    rnd = random.Random(7)
    alpha_p = list("ab.*?[]^-\\/") + ["é", "job", "*."]
    alpha_n = list("ab./-^]") + ["é", "job"]
    n_match = n_bad = 0
    for _ in range(30000):
        pat = "".join(rnd.choice(alpha_p) for _ in range(rnd.randint(0, 7)))
        name = "".join(rnd.choice(alpha_n) for _ in range(rnd.randint(0, 6)))
        try:
            want = 1 if py_oracle.path_match(pat, name) else 0
        except py_oracle.BadPattern:
            want = -1
        got = oracle_lib.path_match(pat, name)
        assert got == want, (pat, name, got, want)
        n_match += want == 1
        n_bad += want == -1
    assert n_match > 500 and n_bad > 500


def test_path_match_agrees_with_fnmatch_on_plain_globs():
    # for patterns made only of literals, '*' and '?' and names without '/', path.Match and
    # Python's fnmatchcase have the same semantics — an implementation independent of both oracles
    rnd = random.Random(3)
    for _ in range(20000):
        pat = "".join(rnd.choice("ab.*?") for _ in range(rnd.randint(0, 6)))
        name = "".join(rnd.choice("ab.") for _ in range(rnd.randint(0, 6)))
        want = 1 if fnmatch.fnmatchcase(name, pat) else 0
        assert oracle_lib.path_match(pat, name) == want, (pat, name)
        assert (1 if py_oracle.path_match(pat, name) else 0) == want, (pat, name)


def test_trim_and_fold_fuzz():
    rnd = random.Random(5)
    chars = [" ", "\t", "\n", " ", " ", "　", "​", "a", "B", "z", "é", "\x1c"]
    for _ in range(5000):
        s = "".join(rnd.choice(chars) for _ in range(rnd.randint(0, 8)))
        assert oracle_lib.trim_space(s).decode() == py_oracle.trim_space(s), repr(s)
        t = "".join(rnd.choice(chars) for _ in range(rnd.randint(0, 4)))
        u = "".join(c.swapcase() if rnd.random() < 0.5 else c for c in t)
        assert oracle_lib.equal_fold(t, u) == py_oracle.equal_fold(t, u), (t, u)
    # Go trims U+0085/U+00A0/U+2000..U+200A/U+3000 but NOT U+200B or U+001C
    assert py_oracle.trim_space("​x​") == "​x​"
    assert py_oracle.trim_space("\x1cx") == "\x1cx"
    assert oracle_lib.trim_space("　x ") == b"x"
