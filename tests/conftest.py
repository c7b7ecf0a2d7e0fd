"""pytest configuration: registers the `gpu` marker (tests that need a real B200) and puts the repo root on
sys.path so `axiom_b200` and `oracle` import without installation."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# One test plays several ranks of the peer-mailbox protocol on separate streams of ONE process; streams that share a
# hardware work queue would serialise a polling kernel ahead of the kernel it is waiting for. 32 queues instead of the
# default 8 (read when the CUDA context is created, i.e. before any test touches the device).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


def _cuda_device_count() -> int:
    try:
        import ctypes
        from axiom_b200 import capi
        n = ctypes.c_int32(0)
        capi.lib().axb_device_count(ctypes.byref(n))
        return n.value
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu():
    """Session fixture for -m gpu tests: binds device 0 and FAILS (never skips) when the CUDA extension or the
    device is missing — a GPU test that silently passes without native code would be a false green."""
    from axiom_b200 import capi
    if _cuda_device_count() < 1:
        pytest.fail("gpu-marked test selected but no CUDA device is visible / libaxb.so failed to load")
    capi.check(capi.lib().axb_set_device(0))
    return capi.lib()
