import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN

sys.path.insert(0, os.path.join(ROOT, 'tests'))


@pytest.fixture(autouse=True)
def _numpy_engine_by_default():
    """Every test starts (and ends) in the NumPy engine, whatever a previous test did."""
    import sqcircuit_b200 as sq
    sq.set_engine('NumPy')
    yield
    sq.set_engine('NumPy')
