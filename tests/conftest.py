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
