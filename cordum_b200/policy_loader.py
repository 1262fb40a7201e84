"""Where the safety policy comes from and how it is hot-reloaded: the safety kernel's loader, restated over the engine.

Mirrors (setup / control code around the hot path):
  policySourceFromEnv        core/controlplane/safetykernel/kernel.go:780-785
  loadPolicyBundle           :787-809   file or URL -> (policy, "<version>:<sha256 hex>" | "<sha256 hex>")
  verifyPolicySignature      :832-849   ed25519 over the raw bytes when SAFETY_POLICY_PUBLIC_KEY is set
  readSignature / decodeKey  :851-881
  policyLoader.Load          :577-588   base bundle + config-service fragments, merged, snapshots combined
  loadFragments              :590-634   fragments in key order, snapshot "cfg:" + sha256(key 0x00 content ...)
  extractPolicyFragment ...  :636-692
  watchPolicy                :485-508   poll every SAFETY_POLICY_RELOAD_INTERVAL (30 s); swap when the snapshot changed
  setPolicy                  :510-521   -> SafetyKernelServer.set_policy -> cordum_policy_load

The swap itself is the engine's: cordum_policy_load compiles the merged document into a fresh table set, uploads it
and bumps the table epoch under the engine lock.  Batches already dispatched finish on the tables they were launched
with; a batch encoded before the swap is refused with CORDUM_E_STALE and encoded again (the front-end does that by
itself), so a request racing a reload is answered under one policy or the other, like the reference's RWMutex read.

The config service (Redis) is out of scope; the loader takes a callable that returns the `bundles` mapping the
reference reads from it (`doc.Data["bundles"]`), which is all loadFragments uses."""
from __future__ import annotations

import base64
import binascii
import hashlib
import os
import threading
import urllib.request
from typing import Callable

from . import policy_io

DEFAULT_RELOAD_INTERVAL_S = 30.0


def policy_source_from_env(path: str) -> str:
    raw = os.environ.get("SAFETY_POLICY_URL", "").strip()
    return raw if raw else (path or "").strip()


def _is_url(source: str) -> bool:
    return source.startswith("http://") or source.startswith("https://")


def read_policy_source(source: str) -> bytes:   # :811-830
    if _is_url(source):
        with urllib.request.urlopen(source, timeout=10) as resp:
            if resp.status < 200 or resp.status >= 300:
                raise IOError("policy fetch status %d" % resp.status)
            return resp.read()
    with open(source, "rb") as f:
        return f.read()


def decode_key(raw: str) -> bytes:   # :870-881: standard base64 first, then hex
    if raw == "":
        raise ValueError("empty key")
    try:
        return base64.b64decode(raw, validate=True)
    except (binascii.Error, ValueError):
        pass
    try:
        return bytes.fromhex(raw)
    except ValueError:
        raise ValueError("invalid key encoding") from None


def read_signature(source: str) -> bytes:   # :851-868
    raw = os.environ.get("SAFETY_POLICY_SIGNATURE", "").strip()
    if raw:
        return decode_key(raw)
    path = os.environ.get("SAFETY_POLICY_SIGNATURE_PATH", "").strip()
    if path:
        with open(path, "rb") as f:
            return f.read()
    if _is_url(source):
        raise ValueError("policy signature required but no signature provided")
    if os.path.exists(source + ".sig"):
        with open(source + ".sig", "rb") as f:
            return f.read()
    raise ValueError("policy signature required but not found")


def verify_policy_signature(data: bytes, source: str) -> None:
    pub_raw = os.environ.get("SAFETY_POLICY_PUBLIC_KEY", "").strip()
    if not pub_raw:
        return
    try:
        pub = decode_key(pub_raw)
    except ValueError as exc:
        raise ValueError("invalid SAFETY_POLICY_PUBLIC_KEY: %s" % exc) from None
    sig = read_signature(source)
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PublicKey

    try:
        Ed25519PublicKey.from_public_bytes(pub).verify(sig, data)
    except (InvalidSignature, ValueError):
        raise ValueError("policy signature verification failed") from None


def load_policy_bundle(source: str):
    """-> (policy | None, snapshot)"""
    if source == "":
        return None, ""
    data = read_policy_source(source)
    verify_policy_signature(data, source)
    policy = policy_io.parse_safety_policy(data)
    digest = hashlib.sha256(data).hexdigest()
    version = (policy or {}).get("version") or ""
    return policy, (str(version) + ":" + digest if version else digest)


def parse_bool(raw: str) -> bool:   # :675-682
    return raw.strip().lower() in ("1", "true", "yes", "y", "on")


def _go_sprint(v) -> str:
    """fmt.Sprint of a decoded JSON value, as far as parseBool can tell the difference."""
    if isinstance(v, float) and v == int(v):
        return str(int(v))
    return "<nil>" if v is None else str(v)


def bundle_enabled(bundle) -> bool:   # :657-673
    if bundle is None or "enabled" not in bundle:
        return True
    v = bundle["enabled"]
    if isinstance(v, bool):
        return v
    return parse_bool(v if isinstance(v, str) else _go_sprint(v))


def extract_policy_fragment(value):   # :636-655 -> (content, ok)
    if isinstance(value, str):
        return value, True
    if isinstance(value, dict):
        if not bundle_enabled(value):
            return "", False
        for key in ("content", "policy", "data"):
            if isinstance(value.get(key), str):
                return value[key], True
    return "", False


def combine_snapshots(base: str, extra: str) -> str:   # :684-692
    if base == "":
        return extra
    if extra == "":
        return base
    return base + "|" + extra


class PolicyLoader:
    """policyLoader.  `bundles`: a callable returning the config document's `bundles` mapping (or None when the
    document / key is absent) - the one thing loadFragments reads from the config service."""

    def __init__(self, source: str = "", bundles: Callable[[], dict | None] | None = None):
        self.source = source
        self.bundles = bundles

    def should_watch(self) -> bool:   # :570-575
        return bool(self.source) or self.bundles is not None

    def load_fragments(self):
        raw = self.bundles() if self.bundles is not None else None
        if not isinstance(raw, dict) or not raw:
            return None, ""
        hasher = hashlib.sha256()
        merged = None
        for key in sorted(raw, key=lambda k: k.encode("utf-8")):     # sort.Strings: byte order
            content, ok = extract_policy_fragment(raw[key])
            if not ok or content.strip() == "":
                continue
            hasher.update(key.encode("utf-8") + b"\x00" + content.encode("utf-8"))
            try:
                policy = policy_io.parse_safety_policy(content)
            except Exception as exc:
                raise ValueError('parse policy fragment "%s": %s' % (key, exc)) from None
            merged = policy_io.merge_policies(merged, policy)
        if merged is None:
            return None, ""
        return merged, "cfg:" + hasher.hexdigest()

    def load(self):
        base, base_snap = load_policy_bundle(self.source)
        frag, frag_snap = self.load_fragments()
        return policy_io.merge_policies(base, frag), combine_snapshots(base_snap, frag_snap)


def reload_interval_from_env() -> float:
    """SAFETY_POLICY_RELOAD_INTERVAL as time.ParseDuration reads the common single-unit forms; anything else -> 30 s."""
    raw = os.environ.get("SAFETY_POLICY_RELOAD_INTERVAL", "")
    units = (("ms", 1e-3), ("us", 1e-6), ("µs", 1e-6), ("ns", 1e-9), ("s", 1.0), ("m", 60.0), ("h", 3600.0))
    for suffix, scale in units:
        if raw.endswith(suffix):
            try:
                d = float(raw[: -len(suffix)]) * scale
            except ValueError:
                break
            return d if d > 0 else DEFAULT_RELOAD_INTERVAL_S
    return DEFAULT_RELOAD_INTERVAL_S


class PolicyWatcher:
    """watchPolicy: reload on a timer; set the policy when the snapshot is non-empty and differs from the one in force.
    A failed reload is logged and skipped (the policy in force stays)."""

    def __init__(self, server, loader: PolicyLoader, interval_s: float | None = None, log=print):
        self.server, self.loader, self.log = server, loader, log
        self.interval_s = reload_interval_from_env() if interval_s is None else interval_s
        self._stop = threading.Event()
        self._thread = None
        self.reloads = 0

    def poll_once(self) -> bool:
        try:
            policy, snapshot = self.loader.load()
        except Exception as exc:
            self.log("safety-kernel: policy reload failed: %s" % exc)
            return False
        if snapshot != "" and snapshot != self.server.engine.current_snapshot():
            self.server.set_policy(policy, snapshot)
            self.reloads += 1
            self.log("safety-kernel: policy snapshot updated %s" % snapshot)
            return True
        return False

    def start(self):
        def run():
            while not self._stop.wait(self.interval_s):
                self.poll_once()
        self._thread = threading.Thread(target=run, name="cordum-policy-watch", daemon=True)
        self._thread.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
