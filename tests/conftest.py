"""pytest configuration: registers the `gpu` marker and builds the CPU oracle on demand.

`-m "not gpu"` : oracle vs the reference's golden vectors, host logic, C-ABI symbol checks (no GPU needed).
`-m gpu`       : parity of the CUDA engine (through the C-ABI) against the oracle, on a real B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    so = os.path.join(ROOT, "oracle", "libbn254_oracle.so")
    src = os.path.join(ROOT, "oracle", "bn254_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    yield
