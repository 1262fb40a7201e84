"""fmt's %q on a string (strconv.Quote), three ways:
   product  cordum_b200/csrc/gostr.hpp go_quote   (via cordum_test_canon kind 2)
   oracle   oracle/oracle.cpp strconv_quote       (structured as go1.24 strconv/quote.go appendEscapedRune)
   python   oracle/py_oracle.py go_quote          (IsPrint from this interpreter's unicodedata 15.0.0, not the shared table)
Reference call sites: core/infra/config/safety_policy.go:235 (tenant deny rule reason), :410,:413 (MCP reasons);
core/controlplane/gateway/policy_bundles.go:1207,1211 (the gateway evaluator's effective-config reasons)."""
import ctypes as C
import random
import sys
import unicodedata

import pytest

import oracle_lib
from cordum_b200 import _lib

sys.path.insert(0, oracle_lib.ORACLE_DIR)
import py_oracle  # noqa: E402


def product_quote(s) -> bytes:
    b = s if isinstance(s, bytes) else s.encode("utf-8", "surrogatepass")
    L = _lib.load()
    buf = C.create_string_buffer(10 * len(b) + 8)
    n = L.cordum_test_canon(2, b, len(b), buf, len(buf))
    return buf.raw[:n]


# Known answers: go1.24 strconv/quote_test.go quotetests (Quote column) and the strconv.Quote documentation examples.
QUOTE_KATS = [
    (b"\a\b\f\r\n\t\v", rb'"\a\b\f\r\n\t\v"'),
    (b"\\", rb'"\\"'),
    (b"abc\xffdef", rb'"abc\xffdef"'),
    ("\u263a".encode(), '"\u263a"'.encode()),
    ("\U0010ffff".encode(), rb'"\U0010ffff"'),
    (b"\x04", rb'"\x04"'),
    ("!\u00a0!\u2000!\u3000!".encode(), rb'"!\u00a0!\u2000!\u3000!"'),   # non-ASCII spaces are not printable
    (b"\x7f", rb'"\x7f"'),
    (b'"Fran & Freddie\'s Diner\t\xe2\x98\xba"', b'"\\"Fran & Freddie\'s Diner\\t\xe2\x98\xba\\""'),   # Quote doc example
    (b"", b'""'),
    (b"job.default", b'"job.default"'),
    ("\u00ad".encode(), rb'"\u00ad"'),          # soft hyphen: Cf
    ("\ufeff".encode(), rb'"\ufeff"'),          # BOM: Cf
    ("\ufffd".encode(), '"\ufffd"'.encode()),   # an encoded U+FFFD is a printable symbol (So), unlike a bad byte
    (b"\xed\xa0\x80", rb'"\xed\xa0\x80"'),      # a UTF-8-encoded surrogate: three bad bytes
    (b"\xc0\x80", rb'"\xc0\x80"'),
    ("\u0085".encode(), rb'"\u0085"'),          # C1 control
    ("\U000e0001".encode(), rb'"\U000e0001"'),  # tag character: Cf
    ("\u0378".encode(), rb'"\u0378"'),          # unassigned
    ("\u65e5\u672c\u8a9e".encode(), '"\u65e5\u672c\u8a9e"'.encode()),
]


@pytest.mark.parametrize("raw,want", QUOTE_KATS)
def test_quote_known_answers(raw, want):
    assert oracle_lib.quote(raw) == want, "oracle"
    assert product_quote(raw) == want, "product"
    assert py_oracle.go_quote(raw).encode("utf-8") == want, "python"


def test_is_print_over_every_code_point():
    """Quote of each single rune: unescaped exactly for categories L, M, N, P, S and U+0020 (unicode.IsPrint)."""
    assert unicodedata.unidata_version == "15.0.0"
    bad = []
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        enc = ch.encode("utf-8")
        want = py_oracle.go_quote(ch).encode("utf-8")
        # assert the shape here too, independent of py_oracle's own branches
        printable = cp == 0x20 or unicodedata.category(ch)[0] in "LMNPS"
        if printable and ch not in '"\\':
            assert want == b'"' + enc + b'"'
        else:
            assert want.startswith(b'"\\')
        if product_quote(enc) != want or oracle_lib.quote(enc) != want:
            bad.append(cp)
    assert not bad, ["U+%04X" % c for c in bad[:20]]


def test_random_byte_strings_agree():
    rng = random.Random(7)
    alphabet = [b"a", b"Z", b" ", b'"', b"\\", b"\n", b"\x00", b"\x1b", b"\x7f", b"\x80", b"\xff", b"\xc3", b"\xc3\xa9", b"\xe2\x82",
                b"\xe2\x82\xac", b"\xf0\x9f\x98\x80", b"\xf0\x9f", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", "\u00a0".encode(),
                "\u200b".encode(), "\u2028".encode(), "\u0085".encode(), "\U000f0000".encode(), "\u0300".encode()]
    for _ in range(20000):
        raw = b"".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 8)))
        a, b, c = product_quote(raw), oracle_lib.quote(raw), py_oracle.go_quote(raw).encode("utf-8")
        assert a == b == c, (raw, a, b, c)
