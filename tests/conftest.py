import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kats():
    import json
    return json.load(open(os.path.join(GOLDEN, "reference_kats.json")))


@pytest.fixture(scope="session")
def testdata_dir():
    return os.path.join(GOLDEN, "testdata")
