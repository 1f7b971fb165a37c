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


def pytest_sessionfinish(session, exitstatus):
    """class-M (marginal) solves are excluded from the bit-exact claim - counted and reported."""
    mod = sys.modules.get('test_cuda_parity')
    counts = getattr(mod, 'CLASS_COUNTS', None) if mod else None
    if not counts:
        return
    import json
    tot = {}
    for c in counts.values():
        for k, v in c.items():
            tot[k] = tot.get(k, 0) + v
    print('\nparity classes (SURVEY Q16) over all oracle comparisons: {}'.format(tot))
    marginal = {t: c for t, c in counts.items() if c.get('M') or c.get('Z')}
    if marginal:
        print('cases with class Z / M solves: {}'.format(json.dumps(marginal)))
    out = os.path.join(REPO, 'gpurun_out')
    if os.path.isdir(out):
        with open(os.path.join(out, 'parity_classes.json'), 'w') as f:
            json.dump({'total': tot, 'by_case': counts}, f, indent=1)
