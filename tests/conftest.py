import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

# Before anything initialises the CUDA driver (torch.cuda.is_available() does): see rayfed_b200/__init__.py
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# keep the adaptive spin short on small CI machines
os.environ.setdefault("RAYFED_B200_SPIN_US", "20")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(autouse=True)
def _clean_fed_state():
    yield
    # a failed in-process test must not poison the next one
    try:
        import rayfed_b200 as fed
        from rayfed_b200._private import global_context, kv

        if global_context.get_global_context() is not None:
            fed.shutdown()
        kv._clear_internal_kv()
    except Exception:  # noqa: BLE001
        pass
