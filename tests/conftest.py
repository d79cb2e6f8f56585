import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU and the built libsrb.so")


def _gpu_unavailable_reason():
    """None when the gpu-marked tests can run here; otherwise why not (they are skipped, not failed)."""
    import torch

    from speech_resynth_b200 import _native

    if not torch.cuda.is_available():
        return "no CUDA device visible"
    if torch.cuda.get_device_capability(0) != (10, 0):
        return "cuda:0 is not an sm_100 (B200) device"
    if not os.path.exists(_native.LIB_PATH):
        return "speech_resynth_b200/libsrb.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    return None


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items:
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def state_dict():
    from speech_resynth_b200 import synthetic

    return synthetic.make_state_dict(seed=0)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
