import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_package():
    """The package directory is literally `cortex.llamacpp_b200` (dot in the name) -> load it by path."""
    name = "cortex_llamacpp_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "cortex.llamacpp_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def b200():
    return load_package()


@pytest.fixture(scope="session")
def ctx(b200):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a GPU; the CUDA path has no fallback"
    c = b200.Context(0)
    yield c
    c.close()
