import pathlib
import sys

import numpy as np
import pytest

REPO = pathlib.Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = REPO / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def unet_golden():
    return dict(np.load(GOLDEN / 'unet_golden.npz'))


@pytest.fixture(scope='session')
def ws_golden():
    return dict(np.load(GOLDEN / 'ws_golden.npz'))


@pytest.fixture(scope='session')
def data_golden():
    return dict(np.load(GOLDEN / 'data_golden.npz'))


@pytest.fixture(scope='session')
def defs_golden():
    return dict(np.load(GOLDEN / 'defs_golden.npz'))


@pytest.fixture(scope='session')
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)


# order of the beta_hat vectors in ws_golden.npz (tests/golden/make_golden.py)
ATTACK_MODES = [(0, False), (0, True), (1, False), (1, True), (-1, False), (-1, True)]
GOLDEN_UNET_CASES = [(0, 24, 40), (1, 32, 48), (2, 64, 64), (2, 40, 72), (3, 64, 64), (4, 64, 96)]
