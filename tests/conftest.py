import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One libmceik_b200 context on cuda:0 for the whole GPU test session (fails loudly without a GPU)."""
    import mceik_b200
    ctx = mceik_b200.Context(0)
    yield ctx
    ctx.close()
