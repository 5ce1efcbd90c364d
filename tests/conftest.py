import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (test infrastructure) and the product library if they are missing or stale."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    import _pkg
    pkg = _pkg.load_package()
    sys.path.insert(0, _pkg.PKG_DIR)
    import importlib.util
    spec = importlib.util.spec_from_file_location("mm_build", os.path.join(_pkg.PKG_DIR, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build_library()
    return pkg


@pytest.fixture(scope="session")
def mm(_built):
    return _built


@pytest.fixture(scope="session")
def assets():
    import scenes
    return scenes.load_assets()


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding
    return oracle_binding
