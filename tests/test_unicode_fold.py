"""strings.EqualFold / strings.ToLower over all of Unicode (go1.24 = Unicode 15.0.0), three ways:
   product  cordum_b200/csrc/gostr.hpp  (canonical forms; via the cordum_test_canon hook)
   oracle   oracle/oracle.cpp           (Go's EqualFold loop over SimpleFold orbits)
   python   oracle/py_oracle.py         (derived from this interpreter's own Unicode 15.0.0 case mappings, NOT from
                                         common/go_unicode_tables.h)
Reference call sites: core/infra/config/safety_policy.go:296-306 (containsString), :209 (normalizeDecision);
core/controlplane/safetykernel/kernel.go:384,388,403; core/controlplane/scheduler/strategy_least_loaded.go:250,256."""
import ctypes as C
import sys
import unicodedata

import numpy as np
import pytest

import oracle_lib
import table_walk
from cordum_b200 import _lib, wire

sys.path.insert(0, oracle_lib.ORACLE_DIR)
import py_oracle  # noqa: E402


def canon(kind: int, s) -> bytes:
    b = s if isinstance(s, bytes) else s.encode("utf-8", "surrogatepass")
    L = _lib.load()
    buf = C.create_string_buffer(4 * len(b) + 8)
    n = L.cordum_test_canon(kind, b, len(b), buf, len(buf))
    return buf.raw[:n]


# known answers: the Go documentation of unicode.SimpleFold / strings.EqualFold and unicode/letter_test.go
# simpleFoldTests ("KkK", "Ssſ", "ρϱΡ", U+0345/Ι/ι/U+1FBE, "İ", "ı" alone, Cherokee upper before lower)
FOLD_KATS = [
    ("Go", "GO", True), ("AB", "ab", True), ("ß", "ss", False), ("K", "k", True), ("K", "K", True),
    ("KELVIN", "Kelvin", True), ("ſ", "S", True), ("ſ", "s", True), ("ς", "Σ", True),
    ("ς", "σ", True), ("İ", "i", False), ("İ", "I", False), ("ı", "I", False), ("ı", "i", False),
    ("Ǆ", "ǅ", True), ("ǅ", "ǆ", True), ("ϴ", "θ", True), ("ϑ", "Θ", True),
    ("ẞ", "ß", True), ("µ", "μ", True), ("µ", "Μ", True), ("Ꭰ", "ꭰ", True),
    ("ͅ", "Ι", True), ("ι", "ι", True), ("ρ", "ϱ", True), ("Ρ", "ϱ", True),
    ("a", "à", False), ("", "", True), ("a", "", False), ("é", "É", True), ("д", "Д", True),
    ("ΐ", "ΐ", False),   # equal only under full case folding (and under Unicode >= 15.1 simple folding)
]


@pytest.mark.parametrize("a,b,want", FOLD_KATS)
def test_equal_fold_known_answers(a, b, want):
    assert oracle_lib.equal_fold(a, b) is want, "oracle"
    assert oracle_lib.equal_fold(b, a) is want, "oracle (swapped)"
    assert py_oracle.equal_fold(a, b) is want, "python"
    assert (canon(0, a) == canon(0, b)) is want, "product"


def test_invalid_utf8_folds_like_go():
    """`for _, r := range s` yields U+FFFD for every invalid byte, so two different invalid bytes are EqualFold."""
    assert oracle_lib.equal_fold(b"\xff", b"\xfe") and oracle_lib.equal_fold(b"a\xffb", b"A\xef\xbf\xbdB")
    assert not oracle_lib.equal_fold(b"\xff", b"\xff\xff")
    assert canon(0, b"\xff") == canon(0, b"\xfe") == b"\xef\xbf\xbd"
    assert canon(0, b"a\xffb") == canon(0, b"A\xef\xbf\xbdB")
    assert canon(1, b"A\xffB") == oracle_lib.to_lower(b"A\xffB") == b"a\xef\xbf\xbdb"


def _all_scalars():
    return [cp for cp in range(0x80, 0x110000) if not (0xD800 <= cp <= 0xDFFF)]


def test_fold_classes_agree_over_all_of_unicode():
    """Partition of every Unicode scalar value into EqualFold classes: product canonical form == python derivation,
    and inside / across classes the oracle's EqualFold loop says the same."""
    cps = _all_scalars()
    text = "".join(map(chr, cps))
    out = canon(0, text).decode("utf-8")
    assert len(out) == len(cps), "the canonical form maps rune to rune"
    prod = {}
    for cp, r in zip(cps, out):
        prod.setdefault(r, []).append(cp)
    py = {}
    for cp in cps:
        py.setdefault(py_oracle._fold_rep(chr(cp)), []).append(cp)
    prod_classes = sorted(tuple(v) for v in prod.values() if len(v) > 1)
    py_classes = sorted(tuple(v) for v in py.values() if len(v) > 1)
    assert prod_classes == py_classes
    assert len(prod_classes) > 1300 and max(len(c) for c in prod_classes) == 4
    # ASCII letters join their classes through the ASCII path (K-sign, long s)
    assert canon(0, "K") == b"k" and canon(0, "ſ") == b"s"
    # oracle: every pair inside a class is EqualFold (both argument orders), neighbours across classes are not
    for cls in prod_classes:
        for a in cls:
            for b in cls:
                assert oracle_lib.equal_fold(chr(a), chr(b)), (hex(a), hex(b))
    reps = sorted(c[0] for c in prod_classes)
    for a, b in zip(reps, reps[1:]):
        if canon(0, chr(a)) != canon(0, chr(b)):
            assert not oracle_lib.equal_fold(chr(a), chr(b)), (hex(a), hex(b))
    # and against ASCII
    for cls in prod_classes:
        for a in cls:
            for ch in "ksiI":
                assert oracle_lib.equal_fold(chr(a), ch) == (canon(0, chr(a)) == canon(0, ch)), (hex(a), ch)


def test_to_lower_agrees_over_all_of_unicode():
    import _sre

    cps = _all_scalars()
    text = "".join(map(chr, cps))
    want = "".join(chr(_sre.unicode_tolower(cp)) for cp in cps)
    assert canon(1, text).decode("utf-8") == want
    assert oracle_lib.to_lower(text).decode("utf-8") == want
    assert py_oracle.to_lower(text) == want
    assert canon(1, "İ") == b"i" and canon(1, "ſ") == "ſ".encode() and canon(1, "K") == b"k"


def test_unicode_version_is_go_1_24s():
    assert unicodedata.unidata_version == "15.0.0"   # go1.24 unicode.Version; the python derivation relies on it


FIELDS = ("decision", "sched_decision", "flags", "route_status", "reason_code", "rule_idx", "worker_slot")


def test_folding_through_the_whole_path():
    """safety_policy.go:296-306: a tenant written with U+212A KELVIN SIGN matches a rule listing "KELVIN" (in Go and here);
    with ASCII-only folding a deny rule would fail open."""
    policy = {"default_tenant": "default", "rules": [
        {"id": "deny-kelvin", "decision": "deny", "reason": "no", "match": {"tenants": ["KELVIN"]}},
        {"id": "deny-sigma", "decision": "deny", "reason": "no", "match": {"capabilities": ["ΣΙΓΜΑΣ"]}},
        {"id": "deny-long-s", "decision": "deny", "reason": "no", "match": {"risk_tags": ["secrets"]}},
        {"id": "dotted", "decision": "throttle", "reason": "no", "match": {"pack_ids": ["İstanbul"]}}]}
    jobs = [{"topic": "job.a", "tenant": "Kelvin"}, {"topic": "job.a", "tenant": "kelvin"},
            {"topic": "job.a", "meta": {"capability": "σιγμας"}},          # final sigma
            {"topic": "job.a", "meta": {"risk_tags": ["ſecretſ"]}},                             # long s: also secrets_present
            {"topic": "job.a", "meta": {"pack_id": "istanbul"}}, {"topic": "job.a", "meta": {"pack_id": "İSTANBUL"}},
            {"topic": "job.a", "tenant": "kelviń"}]
    o = oracle_lib.Oracle(policy, None, [])
    want = o.eval(jobs, wire.MODE_POLICY_ONLY)
    assert [int(x) for x in want["rule_idx"]] == [0, 0, 1, 2, -1, 3, -1]
    h = table_walk.HostHarness(policy, None, [])
    got = h.evaluate(jobs, wire.MODE_POLICY_ONLY)
    for f in FIELDS:
        assert np.array_equal(got[f], want[f]), f
    for j, job in enumerate(jobs):   # the independent python restatement agrees
        rec = py_oracle.process_job(policy, {"topics": {}, "pools": {}}, [], job)
        assert rec["rule_idx"] == int(want["rule_idx"][j]), j
