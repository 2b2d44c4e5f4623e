import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


SIMULATED = os.environ.get("ALOHA_TEST_DEVICE") == "sim"


def pytest_sessionstart(session):
    """ALOHA_TEST_DEVICE=sim: the `gpu` tests run against the engine's host code on the simulated device
    (tests/sim_engine.py) -- a check of the batcher, not of the kernels; the real binding is never loaded."""
    if SIMULATED:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import sim_engine
        session._aloha_sim = sim_engine.simulated()
        session._aloha_sim.__enter__()


def pytest_sessionfinish(session, exitstatus):
    if getattr(session, "_aloha_sim", None) is not None:
        session._aloha_sim.__exit__(None, None, None)


def pytest_collection_modifyitems(config, items):
    if SIMULATED:
        return
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
