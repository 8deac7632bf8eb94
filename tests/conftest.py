import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CPU-side artefacts (oracle, generator) if missing.  The CUDA library is
    built by __graft_entry__.build(); tests that need it fail loudly if it is absent."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "libjoin_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(os.path.join(ROOT, "enclone_b200", "libejoin_synth.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "enclone_b200", "csrc"), "synth"])
    yield
