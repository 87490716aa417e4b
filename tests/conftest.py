import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_meta():
    import json

    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def load_png(name):
    import numpy as np
    from PIL import Image

    return np.array(Image.open(os.path.join(GOLDEN, name)))
