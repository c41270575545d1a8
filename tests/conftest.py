import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """Golden vectors produced by the unmodified reference (tools/gen_golden.py)."""
    import json

    import numpy as np

    z = np.load(os.path.join(ROOT, "tests", "golden", "spatial_golden.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta
