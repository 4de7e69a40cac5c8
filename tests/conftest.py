import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle_models():
    """name -> oracle ProposedEval rebuilt from the seed (cached for the session)."""
    from oracle import cases, proposed
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = cases.build_reference_style_model(proposed.ProposedEval, cases.CODEC_CASES[name])
        return cache[name]

    return get
