import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_engine_factory():
    """Creates engines through the C ABI; fails loudly (no skip) if the CUDA library is missing on a GPU box."""
    from jadespectrogram_b200 import Engine
    made = []

    def make(**kw):
        e = Engine(0, **kw)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()
