import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def act():
    return importlib.import_module("anonymous-credit-tokens_b200")


@pytest.fixture(scope="session")
def octx():
    import corpus
    return corpus.make_ctx(corpus.TEST_PARAMS)


@pytest.fixture(scope="session")
def engine(act, octx):
    """GPU engine with the same params/key as the oracle context."""
    import corpus
    params = act.Params.new(*corpus.TEST_PARAMS)
    assert params.h == octx.h
    key = act.PrivateKey(octx.x, octx.w)
    eng = act.Engine(params, key, device=0)
    yield eng
    eng.close()
