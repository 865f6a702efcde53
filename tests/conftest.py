import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def b200rt():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if not os.path.exists(pkg.LIB_PATH):
        ge.build()
    return pkg


@pytest.fixture(scope="session")
def oracle(b200rt):
    import oracle_binding
    oracle_binding.load()
    return oracle_binding


@pytest.fixture(scope="session")
def fixture_world(b200rt):
    return b200rt.World.fixture()


@pytest.fixture(scope="session")
def gpu_ctx(b200rt, fixture_world):
    ctx = b200rt.Context(0)  # raises B200rtError(ERR_NO_DEVICE) without a GPU: no CPU fallback
    ctx.upload_scene(fixture_world)
    yield ctx
    ctx.close()
