import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    warnings.filterwarnings('ignore', category=SyntaxWarning)


def _ensure_built():
    """Build (or refresh) the CUDA library once per session: nvcc cross-compiles without a GPU."""
    import __graft_entry__ as ge
    ge.build()


_ensure_built()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


@pytest.fixture(scope='session')
def golden():
    return load_golden


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope='session')
def gpu_required():
    if not have_gpu():
        pytest.fail('this test is marked gpu but no CUDA device is visible')
