import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def kat():
    """Reference known-answer fixtures (tests/golden/make_golden.py)."""
    import cv2

    z = np.load(os.path.join(ROOT, "tests/golden/reference_kat.npz"))

    class Kat:
        def frame(self, name):
            return cv2.imdecode(z[name + "_png"], cv2.IMREAD_COLOR)

        def __getitem__(self, k):
            return z[k]

    return Kat()


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, loaded through the python host mirror.  GPU tests only."""
    from fiducials_b200 import _lib

    return _lib.load()
