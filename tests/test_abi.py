"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol
include/cordum_b200.h declares, and refuses to run without a GPU (no CPU fallback).
Also differential tests of the library's own string primitives against the oracle."""
import ctypes as C
import os
import random
import re

import pytest

import oracle_lib
from cordum_b200 import _lib, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cordum_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cordum_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libcordum_b200.so does not export %s" % n
    assert set(names) == set(_lib.API), set(names) ^ set(_lib.API)


def test_struct_sizes_match_header():
    assert C.sizeof(wire.CordumStr) == 8 and C.sizeof(wire.CordumWorkerLoad) == 16
    assert wire.DECISION_DTYPE.itemsize == 16 and wire.LOAD_DTYPE.itemsize == 16
    assert C.sizeof(wire.CordumEnvelopes) == 8 + 8 + 8 + 18 * 8
    assert C.sizeof(wire.CordumWorkers) == 8 + 8 + 8 + 9 * 8


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    L = _lib.load()
    h = C.c_void_p()
    opts = wire.CordumEngineOpts(0, 0, 0, 0)
    rc = L.cordum_engine_create(C.byref(opts), C.byref(h))
    assert rc == -5 and not h.value                      # CORDUM_E_NODEVICE
    assert b"no CPU evaluation path" in L.cordum_last_error()


def test_product_glob_matches_oracle_path_match():
    L = _lib.load()
    rnd = random.Random(11)
    alpha_p = list("ab.*?[]^-\\/") + ["é", "job", "*."]
    alpha_n = list("ab./-^]") + ["é", "job"]
    seen = {1: 0, 0: 0, -1: 0}
    for _ in range(40000):
        pat = "".join(rnd.choice(alpha_p) for _ in range(rnd.randint(0, 7))).encode()
        name = "".join(rnd.choice(alpha_n) for _ in range(rnd.randint(0, 6))).encode()
        want = oracle_lib.path_match(pat, name)
        got = L.cordum_test_glob(pat, len(pat), name, len(name))
        assert got == want, (pat, name, got, want)
        seen[want] += 1
    assert min(seen.values()) > 500
    # invalid UTF-8 in the name and pattern
    for pat, name in [(b"?", b"\xff"), (b"[\xc3\xa9]", b"\xc3\xa9"), (b"a?b", b"a\xe2\x82b"), (b"*\xa9", b"\xc3\xa9")]:
        assert L.cordum_test_glob(pat, len(pat), name, len(name)) == oracle_lib.path_match(pat, name), (pat, name)


def test_product_trim_and_normalize_match_oracle():
    L = _lib.load()
    rnd = random.Random(12)
    pieces = [b" ", b"\t", b"\n", b"\xc2\xa0", b"\xe2\x80\x83", b"\xe3\x80\x80", b"\xe2\x80\x8b", b"a", b"Z", b"\xc3\xa9", b"\xff",
              b"\xe3\x80", b"\x80"]
    for _ in range(20000):
        s = b"".join(rnd.choice(pieces) for _ in range(rnd.randint(0, 7)))
        off, ln = C.c_uint64(), C.c_uint64()
        L.cordum_test_trim(s, len(s), C.byref(off), C.byref(ln))
        assert s[off.value: off.value + ln.value] == oracle_lib.trim_space(s), s
    for raw in ["permit", "block", "require-approval", "allow_with_constraints", "throttle", "", " DENY ", "require_human",
                "maybe", "Allow-With-Constraints", "REQUIRE_APPROVAL"]:
        b = raw.encode()
        assert L.cordum_test_normalize_decision(b, len(b)) == oracle_lib.normalize_decision(raw), raw


def test_product_effective_config_parser_matches_oracle():
    L = _lib.load()
    docs = [b'{"safety":{"denied_topics":["job.deny"]}}', b'{"data":{"safety":{"allowed_topics":["job.*"]}}}', b"", b"[]",
            b"{", b'{"other":1}', b'{"safety":null}', b'{"safety":{"denied_topics":"job.deny"}}',
            b'{"safety":{"denied_topics":"x"},"data":{"safety":{}}}', b'{"safety":{"DENIED_TOPICS":["a","b"]}}',
            b'{"safety":{"pii_detection_enabled":"yes"}}', b'{"safety":{"mcp":{"deny_servers":[1]}}}',
            b'{"safety":{"anomaly_thresholds":{"a":"x"}}}', b'{"safety":{"denied_topics":[null,"a"]}}',
            b'{"safety":{"denied_topics":["a"]},"safety":{"denied_topics":["b","c"]}}', b'null', b'{"data":null}',
            b'{"safety":{"allowed_topics":["x"],"Allowed_Topics":["y","z"]}}']
    for d in docs:
        a, dn = C.c_uint32(), C.c_uint32()
        ok = bool(L.cordum_test_parse_effective(d, len(d), C.byref(a), C.byref(dn)))
        assert (ok, a.value, dn.value) == oracle_lib.parse_effective(d), d


def test_exchange_unique_id_needs_no_gpu():
    """The engine loads libnccl at run time; creating the 128-byte id (rank 0's first step) needs no device.
    When no libnccl can be found the call must fail with an error, not crash."""
    from cordum_b200 import engine
    try:
        a, b = engine.Engine.exchange_unique_id(), engine.Engine.exchange_unique_id()
    except engine.CordumError as e:
        assert "nccl" in str(e).lower()
        return
    assert len(a) == 128 and len(b) == 128 and a != b


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/cordum_b200.h must be consumable by a C compiler (cgo uses one): examples/host_min.c builds as strict C99
    against the header and the shared library, and on a machine without a GPU reports that there is no CPU path."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "host_min")
    libdir = os.path.join(ROOT, "cordum_b200")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_min.c"), "-L" + libdir, "-lcordum_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the example's device path is exercised by hand, not by the CPU suite")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "no CUDA device" in r.stdout, (r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_host_runs_the_device_path(tmp_path):
    """examples/host_min.c - the ABI driven from plain C exactly as a cgo binding would - on the GPU: BASELINE config 1
    (hello-pack: one echo job, one allow rule, a two-worker pool) must come out ALLOW under rule hello-pack-allow and
    routed to the idle worker."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "host_min")
    libdir = os.path.join(ROOT, "cordum_b200")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_min.c"), "-L" + libdir, "-lcordum_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "decision 1 rule hello-pack-allow route_status 1 subject worker.hello-worker-b.jobs" in r.stdout, r.stdout
