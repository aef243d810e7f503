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
    """The package; a fresh checkout has no libact_b200.so yet, so build it first (nvcc cross-compiles sm_100a without a GPU).
    The product itself never builds or falls back: without the library every entry point raises."""
    mod = importlib.import_module("anonymous-credit-tokens_b200")
    if not os.path.exists(mod.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "anonymous-credit-tokens_b200", "csrc"), "-s"])
    return mod


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
