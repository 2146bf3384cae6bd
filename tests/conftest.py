import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the product libraries exist (cheap no-op when up to date)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "pathtrace_rs_b200", "lib", "libpthost.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pathtrace_rs_b200"), "-j4"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    yield
