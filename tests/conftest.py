import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import homerhevc_b200 as hb
        return hb.load_library().hb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: a silent skip would look like a pass
    pass


@pytest.fixture(scope="session")
def ctx():
    import homerhevc_b200 as hb
    c = hb.Context(0)          # raises when the CUDA library or a device is missing: no CPU fallback
    yield c
    c.close()
