import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_half():
    import numpy as np
    import wrmf_cases as wc
    return np.load(os.path.join(wc.GOLDEN, "half_iterations.npz"))


@pytest.fixture(scope="session")
def golden_traces():
    import numpy as np
    import wrmf_cases as wc
    return np.load(os.path.join(wc.GOLDEN, "als_traces.npz"))


@pytest.fixture(scope="session")
def cases():
    import wrmf_cases as wc
    return wc.half_iteration_cases()


@pytest.fixture(scope="session")
def golden_bias():
    import numpy as np
    import wrmf_cases as wc
    return np.load(os.path.join(wc.GOLDEN, "bias_half_iterations.npz"))


@pytest.fixture(scope="session")
def bias_cases():
    import wrmf_cases as wc
    return wc.bias_cases()
