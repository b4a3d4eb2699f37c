import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "also_gpu: CPU-only test that is ALSO selected by `-m gpu`, so that the GPU run's record "
                            "carries the oracle <-> compiled-reference pin (oracle/_ref travels to the GPU box prebuilt)")


@pytest.hookimpl(tryfirst=True)
def pytest_collection_modifyitems(config, items):
    expr = (config.getoption("markexpr", "") or "").strip()
    if expr == "gpu":  # the driver's GPU run: pull the also_gpu tests in
        for it in items:
            if it.get_closest_marker("also_gpu"):
                it.add_marker(pytest.mark.gpu)


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
