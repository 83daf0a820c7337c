import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fuzz_cases():
    with gzip.open(os.path.join(GOLDEN, "fuzz.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def medium_cases():
    with open(os.path.join(GOLDEN, "medium.json")) as f:
        return json.load(f)
