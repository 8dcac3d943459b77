import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def trusted_setup_bytes():
    """(setup_G1 compressed (4096,48), setup_G1_lagrange compressed (4096,48)) from eth/trusted_setup.json"""
    import numpy as np
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "trusted_setup_g1.bin"), dtype=np.uint8)
    raw = raw.reshape(2, 4096, 48)
    return raw[0], raw[1]
