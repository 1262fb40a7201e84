"""common/mini_json.hpp parses the policy / routing / effective-config documents for BOTH the product and the C++
oracle, so a parser bug would be invisible to every product-vs-oracle comparison.  This test pins it against an
independent parser (Python's json, strict) on 100,000 seeded documents: valid documents with exotic escapes / numbers /
whitespace / duplicate keys, byte-level corruptions of them, and invalid UTF-8 (encoding/json coerces string contents
to well-formed UTF-8, one U+FFFD per bad byte)."""
import ctypes as C
import json
import random

from cordum_b200 import _lib

N_DOCS = 100_000


def canon(doc: bytes):
    L = _lib.load()
    buf = C.create_string_buffer(8 * len(doc) + 64)
    n = L.cordum_test_json_canon(doc, len(doc), buf, len(buf))
    return None if n < 0 else buf.raw[:n]


def go_coerce_utf8(b: bytes) -> str:
    """utf8.DecodeRune over the bytes: a byte that does not start a valid sequence -> U+FFFD, width 1."""
    out, i, n = [], 0, len(b)
    while i < n:
        c = b[i]
        w = 0
        if c < 0x80:
            w = 1
        elif 0xC2 <= c <= 0xF4:
            need = 3 if c >= 0xF0 else 2 if c >= 0xE0 else 1   # continuation bytes
            if i + need < n:
                lo, hi = 0x80, 0xBF
                if c == 0xE0:
                    lo = 0xA0
                elif c == 0xED:
                    hi = 0x9F
                elif c == 0xF0:
                    lo = 0x90
                elif c == 0xF4:
                    hi = 0x8F
                if lo <= b[i + 1] <= hi and all((b[i + k] & 0xC0) == 0x80 for k in range(2, need + 1)):
                    w = need + 1
        if w:
            out.append(b[i:i + w].decode("utf-8"))
            i += w
        else:
            out.append("�")
            i += 1
    return "".join(out)


def _no_constants(name):
    raise ValueError("NaN / Infinity are not JSON")


def py_parse(doc: bytes):
    """(ok, value) by Python's strict parser on the Go-coerced text (outside strings U+FFFD is a syntax error for both)."""
    try:
        return True, json.loads(go_coerce_utf8(doc), parse_constant=_no_constants)
    except (ValueError, RecursionError):
        return False, None


def fix_surrogates(v):
    """Python keeps a lone \\uD800 escape as a lone surrogate; encoding/json (and mini_json) make it U+FFFD."""
    if isinstance(v, str):
        return "".join("�" if 0xD800 <= ord(c) <= 0xDFFF else c for c in v)
    if isinstance(v, list):
        return [fix_surrogates(x) for x in v]
    if isinstance(v, dict):
        return {fix_surrogates(k): fix_surrogates(x) for k, x in v.items()}
    return v


def same(a, b) -> bool:
    if isinstance(a, bool) or isinstance(b, bool) or a is None or b is None:
        return type(a) is type(b) and a == b
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        if isinstance(a, int) and isinstance(b, int):
            return a == b
        try:
            return float(a) == float(b)   # beyond int64 mini_json keeps the double (as encoding/json's float64 would)
        except OverflowError:
            return False
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
    return False


WS = [" ", "\t", "\n", "\r", ""]
STR_ATOMS = ["a", "Z", " ", "job.*", "\\n", "\\\"", "\\\\", "\\/", "\\b", "\\f", "\\r", "\\t", "\\u0041", "\\u00e9", "\\u212a",
             "\\ud83d\\ude00", "\\ud800", "\\udc00x", "\\uD834\\uDD1E", "é", "K", "😀", "\x7f", "{", "]", ",", ":"]
NUMS = ["0", "-0", "1", "-1", "12", "9223372036854775807", "-9223372036854775808", "9223372036854775808",
        "123456789012345678901234567890", "0.5", "-0.25", "1e3", "1E-3", "1.5e+10", "2e308", "4.9e-324", "1e-400", "0e0", "100"]


def gen_string(rng) -> str:
    return '"' + "".join(rng.choice(STR_ATOMS) for _ in range(rng.randrange(0, 6))) + '"'


def gen_value(rng, depth=0) -> str:
    r = rng.random()
    w = lambda: rng.choice(WS)
    if depth > 5 or r < 0.35:
        k = rng.randrange(6)
        return [gen_string(rng), rng.choice(NUMS), "true", "false", "null", gen_string(rng)][k]
    if r < 0.65:
        items = [gen_value(rng, depth + 1) for _ in range(rng.randrange(0, 4))]
        return "[" + w() + ("," + w()).join(items) + w() + "]"
    keys = ["safety", "data", "allowed_topics", "denied_topics", "mcp", "k", "k", gen_string(rng)[1:-1]]
    members = []
    for _ in range(rng.randrange(0, 4)):
        members.append('"%s"%s:%s%s' % (rng.choice(keys), w(), w(), gen_value(rng, depth + 1)))
    return "{" + w() + ("," + w()).join(members) + w() + "}"


BAD_BYTES = [b"\xff", b"\xc0\x80", b"\xe2\x82", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\x80", b"\xc3", b"\x00", b"\x1f"]
SNIPPETS = [b"NaN", b"Infinity", b"-", b"01", b"1.", b".5", b"1e", b"+1", b"tru", b"nul", b"'a'", b",", b":", b"\"", b"\\", b"\\x",
            b"\\u12", b"[", b"]", b"{", b"}", b"/*c*/", b" ", b"\n", b"1 2", b"\"a\" \"b\""]


def corrupt(rng, doc: bytes) -> bytes:
    b = bytearray(doc)
    for _ in range(rng.randrange(1, 3)):
        k = rng.randrange(5)
        pos = rng.randrange(len(b) + 1)
        if k == 0 and b:
            del b[rng.randrange(len(b))]
        elif k == 1:
            b[pos:pos] = rng.choice(SNIPPETS)
        elif k == 2:
            b[pos:pos] = rng.choice(BAD_BYTES)
        elif k == 3 and b:
            b[rng.randrange(len(b))] = rng.randrange(256)
        else:
            b = b[:pos]
    return bytes(b)


def test_mini_json_against_python_json_on_100k_documents():
    rng = random.Random(20260921)
    n_ok = n_bad = 0
    for i in range(N_DOCS):
        doc = (rng.choice(WS) + gen_value(rng) + rng.choice(WS)).encode("utf-8")
        if i % 3:
            doc = corrupt(rng, doc)
        ok, want = py_parse(doc)
        got = canon(doc)
        assert (got is not None) == ok, "accept/reject differs on %r: python %s, mini_json %s" % (doc, ok, got is not None)
        if not ok:
            n_bad += 1
            continue
        n_ok += 1
        back = json.loads(got.decode("utf-8"))
        assert same(back, fix_surrogates(want)), "value differs on %r: %r vs %r" % (doc, back, want)
    assert n_ok > 30_000 and n_bad > 20_000, (n_ok, n_bad)   # both sides of the accept/reject line are exercised


def test_deep_nesting_and_duplicates():
    assert canon(b"[" * 200 + b"]" * 200) is not None
    assert canon(b"[" * 300 + b"]" * 300) is None            # the reader's own depth limit (documents here are shallow)
    assert json.loads(canon(b'{"a":1,"a":2}'))["a"] == 2      # last duplicate wins, as encoding/json
    assert canon(b'"\xff\xfe"') == '"��"'.encode()
    assert canon(b'"\xe2\x82"') == '"��"'.encode()  # one replacement per bad byte (utf8.DecodeRune)
