import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native artefacts exist (no-op when already built in-tree)."""
    lib = os.path.join(ROOT, "relion_b200", "librelion_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    shim = os.path.join(ROOT, "tests", "cpp", "estep_multi_gpu")
    if not (os.path.exists(lib) and os.path.exists(orc) and os.path.exists(shim)):
        import __graft_entry__ as g
        g.build()
    yield


@pytest.fixture(scope="session")
def device():
    from relion_b200.estep import MlDeviceBundle
    d = MlDeviceBundle(0)
    yield d
    d.close()
