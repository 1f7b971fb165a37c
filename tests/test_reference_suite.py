"""The reference's OWN pytest suite (tests/python/test_stiffness_checker.py of pyconmech 0.6.0:
solve consistency, nodal loads, neighbour queries, nodal equilibrium, self weight, UDL + gravity, the
MGZ 5.8 analytical beam) run against ``checker_engine="cuda"`` through the run-time switch of
``conmech_b200.pyconmech_adapter`` - the executed form of the INTEGRATION.md section 1 stub.

The reference install and a copy of its tests live under the git-ignored ``baseline/_ref`` (made by
``oracle.ref_bridge.install()`` in the authoring container, shipped to the GPU box by gpurun); the
test skips when that directory is absent.  The reference's test file is used UNMODIFIED; only the
fixtures its conftest provides are restated here, because that conftest restricts ``--engine`` to
{cpp, numpy} (tests/conftest.py:6-12).
"""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from oracle import ref_bridge

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFTEST = '''
import os, sys
import pytest
sys.path.insert(0, {repo!r})
from oracle import ref_bridge
ref_bridge.install_stubs()
sys.path.insert(0, ref_bridge.REF_DIR)
ENGINE = os.environ.get('CONMECH_TEST_ENGINE', 'cuda')
if ENGINE == 'cuda':
    from conmech_b200 import pyconmech_adapter
    pyconmech_adapter.install()

@pytest.fixture
def test_data_dir():
    return {data!r}

@pytest.fixture
def n_attempts():
    return 10

@pytest.fixture
def viewer():
    return False

@pytest.fixture
def debug():
    return False

@pytest.fixture
def engine():
    return ENGINE
'''


def run_reference_suite(engine, tmp_path):
    """Runs the reference's test_stiffness_checker.py in a child pytest; returns (rc, tail of output)."""
    src = os.path.join(ref_bridge.REF_TESTS, 'python', 'test_stiffness_checker.py')
    data = os.path.join(ref_bridge.REF_TESTS, 'test_data')
    work = os.path.join(str(tmp_path), 'ref_suite_' + engine)
    os.makedirs(os.path.join(work, 'python'), exist_ok=True)
    shutil.copy(src, os.path.join(work, 'python', 'test_stiffness_checker.py'))
    os.symlink(data, os.path.join(work, 'test_data'))          # the suite locates its data relative to its own file
    with open(os.path.join(work, 'conftest.py'), 'w') as f:
        f.write(textwrap.dedent(CONFTEST.format(repo=REPO, data=data)))
    env = dict(os.environ, CONMECH_TEST_ENGINE=engine)
    r = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-x', '-p', 'no:cacheprovider', '--rootdir', work,
                        '-W', 'ignore', os.path.join(work, 'python', 'test_stiffness_checker.py')],
                       cwd=work, env=env, capture_output=True, text=True, timeout=1500)
    return r.returncode, (r.stdout + r.stderr)[-6000:]


needs_ref = pytest.mark.skipif(not (ref_bridge.available() and os.path.isdir(ref_bridge.REF_TESTS)),
                               reason='reference install baseline/_ref absent (oracle.ref_bridge.install())')


@needs_ref
def test_harness_runs_the_reference_suite_on_its_numpy_engine(tmp_path):
    """CPU: the harness itself - the reference's suite passes on the reference's own engine."""
    rc, tail = run_reference_suite('numpy', tmp_path)
    assert rc == 0, tail
    assert ' passed' in tail


@needs_ref
@pytest.mark.gpu
def test_reference_suite_passes_on_the_cuda_engine(tmp_path):
    rc, tail = run_reference_suite('cuda', tmp_path)
    assert rc == 0, tail
    assert ' passed' in tail
    print(tail[-400:])
