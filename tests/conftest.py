import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer-running CPU test")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, as __graft_entry__.build() does
    import subprocess
    for lib, src in ((os.path.join(ROOT, "cordum_b200", "libcordum_b200.so"), os.path.join(ROOT, "cordum_b200", "csrc")),
                     (os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "oracle"))):
        if not os.path.exists(lib):
            subprocess.run(["make", "-C", src], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib


def _have_cuda_device() -> bool:
    """True when cordum_engine_create finds a device (it answers CORDUM_E_NODEVICE without one: there is no CPU path)."""
    try:
        import ctypes as C

        from cordum_b200 import _lib, wire

        L = _lib.load()
        h = C.c_void_p()
        rc = L.cordum_engine_create(C.byref(wire.CordumEngineOpts(0, 0, 0, 0)), C.byref(h))
        if rc == 0:
            L.cordum_engine_destroy(h)
        return rc == 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `pytest tests` on a machine without a GPU: the gpu-marked tests are skipped, not errors
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _have_cuda_device():
        skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU path; run with -m gpu on a B200)")
        for it in gpu_items:
            it.add_marker(skip)
