import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the C oracle must exist before any test imports them."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("kmcuda_b200_build", os.path.join(ROOT, "kmcuda_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from oracle import oracle as O
    O.build()
    yield
