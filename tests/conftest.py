import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # build the C-ABI library and the host packer before any test module imports the package
    # (a fresh clone has no binaries; with everything up to date this is a few file-time checks)
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(REPO, 'tests', 'golden')


@pytest.fixture(scope='session')
def models_dir():
    return os.path.join(REPO, 'tests', 'golden', 'models')
