"""Access to the UNMODIFIED reference (pyconmech 0.6.0) for tests and for the CPU arm of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY - nothing under conmech_b200/ imports this module.

The reference is a pure-Python package.  It is installed once, in the authoring container, into the
git-ignored directory ``baseline/_ref`` (``install()`` below: ``pip install --no-index --no-deps
--target baseline/_ref`` from a copy of /root/reference, plus a copy of the reference's own
``tests/`` tree under ``baseline/_ref/_tests`` for tests/test_reference_suite.py); ``gpurun`` ships that
directory to the GPU box, where /root/reference does not exist.  The reference imports ``termcolor``
(and its tests ``matplotlib``), which this image lacks: inert stand-ins are registered in
``sys.modules`` - the reference's own files are not touched.
"""
import json
import os
import shutil
import subprocess
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, 'baseline', '_ref')
REF_TESTS = os.path.join(REF_DIR, '_tests')
REF_SRC = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'pyconmech', 'frame_analysis'))


def install(force=False):
    """Populate baseline/_ref from /root/reference (authoring container only).  Returns True when the
    reference is available afterwards."""
    if available() and not force and os.path.isdir(REF_TESTS):
        return True
    if not os.path.isdir(REF_SRC):
        return available()
    import tempfile
    tmp = tempfile.mkdtemp(prefix='refcopy_')
    try:
        src = os.path.join(tmp, 'reference')          # the build writes into the tree; /root/reference is read-only
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns('.git'))
        os.makedirs(REF_DIR, exist_ok=True)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--upgrade',
               '--find-links', '/opt/wheelhouse', '--target', REF_DIR, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
            return available()
        if os.path.isdir(REF_TESTS):
            shutil.rmtree(REF_TESTS)
        shutil.copytree(os.path.join(REF_SRC, 'tests'), REF_TESTS, ignore=shutil.ignore_patterns('cpp', '__pycache__'))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return available()


def install_stubs():
    """Inert ``termcolor`` / ``matplotlib.pyplot`` modules when the real ones are missing."""
    try:
        import termcolor  # noqa: F401
    except ImportError:
        tc = types.ModuleType('termcolor')
        tc.cprint = lambda *a, **k: None
        tc.colored = lambda s, *a, **k: s
        sys.modules['termcolor'] = tc
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules.setdefault('matplotlib', mpl)
        sys.modules.setdefault('matplotlib.pyplot', plt)


def import_reference():
    """(StiffnessChecker, io_base module) of the reference; raises ImportError when it is not installed."""
    if not available():
        raise ImportError('reference not installed under baseline/_ref (oracle.ref_bridge.install())')
    install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')                  # SyntaxWarning: invalid escape sequences in its docstrings
        from pyconmech.frame_analysis.stiffness_checker import StiffnessChecker
        from pyconmech.frame_analysis import io_base
    return StiffnessChecker, io_base


def ref_model_from_ours(rio, model):
    """One of our Models as a reference Model, through the JSON data layout both share."""
    return rio.Model.from_data(json.loads(json.dumps(model.to_data())))


def ref_loadcase_from_ours(rio, lc):
    pls = [rio.PointLoad(pl.force, pl.moment, pl.node_ind) for pl in lc.point_loads]
    uls = [rio.UniformlyDistLoad(ul.q, ul.load, ul.elem_tags) for ul in lc.uniform_element_loads]
    g = rio.GravityLoad(lc.gravity_load.force) if lc.gravity_load is not None else None
    return rio.LoadCase(pls, uls, g)


def make_ref_checker(model, loadcase, trans_tol=1e-3):
    """The reference facade on the numpy engine for one of our models / load cases."""
    import contextlib
    import io
    RefChecker, rio = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        sc = RefChecker(ref_model_from_ours(rio, model), checker_engine='numpy')
        sc.set_loads(ref_loadcase_from_ours(rio, loadcase))
        sc.set_nodal_displacement_tol(trans_tol=trans_tol)
    return sc
