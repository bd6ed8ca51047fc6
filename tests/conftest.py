import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def synth_small():
    """3 proteomes x 120 genes, seeded."""
    from spacedust_b200 import synth
    return synth.generate(3, 120, seed=7)


@pytest.fixture(scope="session")
def gpu_ctx():
    import spacedust_b200 as sd
    ctx = sd.Context(0)
    yield ctx
    ctx.close()
