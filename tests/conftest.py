import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def configs():
    return load_golden("configs.json")


@pytest.fixture(scope="session")
def kats():
    return load_golden("kats.json")


@pytest.fixture(scope="session")
def pairs():
    return load_golden("pairs.json")


@pytest.fixture(scope="session")
def toml_golden():
    return load_golden("toml_golden.json")
