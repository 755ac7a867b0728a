import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build in-tree artefacts when they are missing (CPU container: nvcc cross-compiles)."""
    import gt16_testlib as T
    need = [T.LIB, T.CLI, T.SYNTH, T.EMU_LIB, T.ORACLE_LIB]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()
    yield


@pytest.fixture(scope="session")
def gpu_caller():
    """One context (tables + decision bitmaps built once) shared by the GPU tests."""
    import gt16_testlib as T
    c = T.Caller(0)
    yield c
    c.close()


def pytest_sessionfinish(session, exitstatus):
    import producer_fixture
    producer_fixture.cleanup()
