import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def models():
    """(full, small) synthetic FRC models (the real blob is missing from the reference mount)."""
    from oracle import synth_model
    return synth_model.ensure_models()


@pytest.fixture(scope="session")
def tod():
    import tod_b200
    return tod_b200
