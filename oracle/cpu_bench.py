"""CPU timing of the reference's path (test/bench infrastructure only) on the host cores over a
bounded sample of a workload.  Used by bench.py for the ``cpu_baseline`` object, the CPU rates of
``other_configs`` and for ``--impl reference``.

kind = "reference": the UNMODIFIED reference, ``StiffnessChecker(model, checker_engine="numpy")
.solve(ids)`` (numpy_stiffness.py:228-397), imported from the git-ignored install under
``baseline/_ref`` that ``gpurun`` ships to the GPU box (oracle/ref_bridge.py).
kind = "port": the oracle restatement (oracle/frame_oracle.py) - used when the reference install is
absent, and always reported beside the reference.  The port vectorises the reference's
interpreted triple loop in ``assemble_global_stiffness_matrix`` (numpy_stiffness_assembly.py:
379-387) and its per-DOF Python loops; everything else (scipy csc products, SuperLU with the same
options) is the same call sequence, so the port is about 2x FASTER than the reference.

A workload is a dict (the same definitions bench.py drives the GPU with):
    {'model': ('grid', nx, ny, nz, joints) | ('json', file name under tests/golden/models),
     'loads': 'gravity' | 'config4',
     'subsets': 'connected' | 'connected:lo:hi' | 'heightfield' | 'full' | 'bfs_prefixes',
     'trans_tol': float}
One "solve" = one (subset, load case); the reference solves one load case per call with a fresh
checker per load case.
"""
import os
import time

_STATE = {}
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIG3 = {'model': ('grid', 7, 7, 8, False), 'loads': 'gravity', 'subsets': 'connected', 'trans_tol': 1e-3}


def reference_available():
    from oracle import ref_bridge
    return ref_bridge.available()


def build_workload(spec):
    """(model, [LoadCase], subset function seed -> ids)"""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad, Model
    m = spec['model']
    if m[0] == 'grid':
        model = syn.grid_frame(m[1], m[2], m[3], joints=bool(m[4]))
    else:
        model = Model.from_json(os.path.join(REPO, 'tests', 'golden', 'models', m[1]))
    if spec['loads'] == 'config4':
        lcs = syn.config4_loadcases(model)
    elif spec['loads'].startswith('json:'):
        from conmech_b200 import LoadCase as LC
        lcs = [LC.from_json(os.path.join(REPO, 'tests', 'golden', 'models', spec['loads'][5:]))]
    else:
        lcs = [LoadCase(gravity_load=GravityLoad([0, 0, -1]))]
    kind = spec['subsets']
    if kind.startswith('connected'):
        parts = kind.split(':')
        lo, hi = (float(parts[1]), float(parts[2])) if len(parts) == 3 else (0.2, 0.9)
        sub = lambda s: syn.connected_subset(model, s, lo=lo, hi=hi)
    elif kind == 'heightfield':
        sub = lambda s: syn.heightfield_subset(model, m[1], m[2], m[3], seed=s)
    elif kind == 'bfs_prefixes':
        order = syn.bfs_sequence(model)
        sub = lambda s: order[:(s % len(order)) + 1]
    elif kind.startswith('bfs_prefixes_spread:'):          # n prefixes evenly spread over the sequence
        order = syn.bfs_sequence(model)
        n = int(kind.split(':')[1])
        sub = lambda s: order[:max(1, ((s % n) + 1) * len(order) // n)]
    else:
        sub = lambda s: []
    return model, lcs, sub


def _init_worker(spec, kind='port'):
    """One checker per (worker process, load case): the oracle port, or the reference facade on its
    numpy engine."""
    import warnings
    warnings.simplefilter('ignore')
    try:                                  # one BLAS/OpenMP thread per worker: the pool owns the cores
        from threadpoolctl import threadpool_limits
        _STATE['tp'] = threadpool_limits(1)
    except Exception:
        pass
    model, lcs, sub = build_workload(spec)
    solvers = []
    checkers = []
    for lc in lcs:
        if kind == 'reference':
            from oracle import ref_bridge
            sc = ref_bridge.make_ref_checker(model, lc, spec['trans_tol'])

            def solve(ids, sc=sc):                              # the reference's public call, stock code path
                try:
                    return bool(sc.solve(ids))
                except TypeError:                               # bare False from compute_dof_permutation
                    return False
        else:
            from oracle.frame_oracle import FrameOracle
            orc = FrameOracle(model)
            orc.set_loads(lc)
            orc.set_nodal_displacement_tol(spec['trans_tol'])

            def solve(ids, orc=orc):
                return orc.solve(ids).success
        solvers.append(solve)
        checkers.append(sc if kind == 'reference' else orc)
    _STATE['solvers'] = solvers
    _STATE['checkers'] = checkers
    _STATE['kind'] = kind
    _STATE['shape'] = (len(model.nodes), len(model.elements))
    _STATE['sub'] = sub


def _solve_seeds(seeds):
    import contextlib
    import io
    solvers, sub = _STATE['solvers'], _STATE['sub']
    subsets = [sub(s) for s in seeds]                              # not timed
    n_ok = 0
    with contextlib.redirect_stdout(io.StringIO()):                # the reference prints per solve
        t0 = time.perf_counter()
        for ids in subsets:
            for solve in solvers:
                n_ok += 1 if solve(ids) else 0
        dt = time.perf_counter() - t0
    return dt, len(seeds) * len(solvers), n_ok


def _parity_seeds(seeds):
    """Dense results (first load case) of the subsets of ``seeds`` from this worker's checker - the reference's
    engine (``get_solved_results`` rows scattered by node / element id, SURVEY section 8c: never compare on row
    order) or the port.  Rows of absent elements are NaN."""
    import contextlib
    import io
    import numpy as np
    chk, kind, sub = _STATE['checkers'][0], _STATE['kind'], _STATE['sub']
    nV, nE = _STATE['shape']
    n = len(seeds)
    out = {'flag': np.zeros(n, np.uint8), 'has': np.zeros(n, np.uint8), 'max_trans': np.full(n, np.nan),
           'compliance': np.full(n, np.nan), 'U': np.zeros((n, 6 * nV)), 'R': np.zeros((n, 6 * nV)),
           'eR': np.full((n, nE, 12), np.nan)}
    with contextlib.redirect_stdout(io.StringIO()):
        for k, s in enumerate(seeds):
            ids = sub(s)
            if kind == 'reference':
                eng = chk._sc_ins
                try:
                    flag = bool(chk.solve(ids))
                except TypeError:
                    flag = False
                    eng._has_stored_deformation = False
                has = bool(chk.has_stored_result())
                if has:
                    _, nD, fR, eR = eng.get_solved_results()
                    out['max_trans'][k] = eng.get_max_nodal_deformation()[0]
                    out['compliance'][k] = eng.get_compliance()
                    for row in nD:
                        out['U'][k, 6 * int(row[0]):6 * int(row[0]) + 6] = row[1:]
                    for row in fR:
                        out['R'][k, 6 * int(row[0]):6 * int(row[0]) + 6] = row[1:]
                    for row in eR:
                        out['eR'][k, int(row[0])] = row[1:]
            else:
                o = chk.solve(ids)
                flag, has = bool(o.success), bool(o.has_result)
                if has:
                    out['max_trans'][k], out['compliance'][k] = o.max_trans, o.compliance
                    out['U'][k], out['R'][k] = o.U, o.R
                    e_ids = ids if len(ids) else list(range(nE))
                    out['eR'][k, e_ids] = np.asarray(o.eR)[e_ids]
            out['flag'][k], out['has'][k] = flag, has
    return out


def parity_sample(spec, n, path, kind=None):
    """Solve the first ``n`` subsets of a workload with the reference (the port when it is not installed) on all
    host cores and store the dense results in ``path`` (.npz): the parity set bench.py checks the GPU results of the
    SAME run against.  Returns the kind used."""
    import multiprocessing as mp
    import numpy as np
    kind = kind or ('reference' if reference_available() else 'port')
    seeds = list(range(n))
    n_procs = max(1, min(os.cpu_count() or 1, n))
    chunks = [seeds[i::n_procs] for i in range(n_procs)]
    ctx = mp.get_context('spawn')
    with ctx.Pool(n_procs, initializer=_init_worker, initargs=(spec, kind)) as pool:
        res = pool.map(_parity_seeds, chunks)
    merged = {}
    for key in res[0]:
        arr = np.zeros((n,) + res[0][key].shape[1:], dtype=res[0][key].dtype)
        for i, r in enumerate(res):
            arr[i::n_procs] = r[key]
        merged[key] = arr
    np.savez(path, kind=np.array(kind), **merged)
    return kind


def time_workload(spec, seeds, n_procs=1, kind='port'):
    """Solve the subsets of ``seeds`` (every load case each) on ``n_procs`` processes.
    Returns (solves_per_second, wall_seconds_of_the_solve_phase, n_true_flags)."""
    seeds = list(seeds)
    if n_procs <= 1:
        _init_worker(spec, kind)
        _solve_seeds(seeds[:1])
        dt, n, ok = _solve_seeds(seeds)
        return n / dt, dt, ok
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    chunks = [seeds[i::n_procs] for i in range(n_procs)]
    chunks = [c for c in chunks if c]
    with ctx.Pool(len(chunks), initializer=_init_worker, initargs=(spec, kind)) as pool:
        pool.map(_solve_seeds, [[s] for s in seeds[:len(chunks)]])          # warm-up: imports, first solve
        t0 = time.perf_counter()
        res = pool.map(_solve_seeds, chunks)
        wall = time.perf_counter() - t0
    n = sum(r[1] for r in res)
    ok = sum(r[2] for r in res)
    return n / wall, wall, ok


def time_oracle(grid, seeds, n_procs=1, trans_tol=1e-3, kind='port'):
    spec = dict(CONFIG3, model=('grid', grid[0], grid[1], grid[2], False), trans_tol=trans_tol)
    return time_workload(spec, seeds, n_procs, kind)


def cpu_model_name():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def _rate_record(spec, n_sample, n_procs, kind, with_single=False):
    rate, wall, n_true = time_workload(spec, range(n_sample), n_procs=n_procs, kind=kind)
    rec = {'value': rate, 'unit': 'solves/s', 'cores': n_procs, 'kind': kind,
           'sample': '{} subsets on {} process(es), {:.1f} s'.format(n_sample, n_procs, wall), 'flags_true': n_true}
    if with_single:
        rate1, _, _ = time_workload(spec, range(min(n_sample, 24)), n_procs=1, kind=kind)
        rec['single_core_value'] = rate1
    return rec


def baseline_record(grid, trans_tol, batch, sample=0):
    """The ``cpu_baseline`` object of bench.py: the reference (when installed, else the port) on all
    host cores over a bounded sample of the config-3 workload; the other implementation's rate beside it."""
    n_procs = os.cpu_count() or 1
    have_ref = reference_available()
    kind = 'reference' if have_ref else 'port'
    spec = dict(CONFIG3, model=('grid', grid[0], grid[1], grid[2], False), trans_tol=trans_tol)
    n_sample = sample or min(batch, (40 if have_ref else 100) * n_procs)          # ~5 s per process either way
    rec = _rate_record(spec, n_sample, n_procs, kind, with_single=True)
    rec['sample'] = 'first {}; single process: {:.1f} solves/s; {}'.format(rec['sample'], rec['single_core_value'], cpu_model_name())
    if have_ref:
        rec['what'] = 'StiffnessChecker(checker_engine="numpy").solve of the unmodified reference (baseline/_ref)'
        prate, _, _ = time_workload(spec, range(min(batch, 100 * n_procs)), n_procs=n_procs, kind='port')
        rec['port_value'] = prate
        rec['port_note'] = 'oracle/frame_oracle.py (vectorised restatement) on the same cores: faster than the reference'
    else:
        rec['what'] = 'oracle port (reference install baseline/_ref absent)'
    return rec


# the other BASELINE.json configurations, as bench.py's `other_configs` runs them on the GPU
OTHER_CONFIGS = {
    'config1_tower_full_solve': {'model': ('json', 'tower_3D.json'), 'loads': 'gravity', 'subsets': 'full', 'trans_tol': 1e-3},
    'config2_tower_bfs_prefixes': {'model': ('json', 'tower_3D.json'), 'loads': 'gravity', 'subsets': 'bfs_prefixes', 'trans_tol': 1e-3},
    'config4_joints_4_loadcases': {'model': ('grid', 7, 7, 8, True), 'loads': 'config4', 'subsets': 'heightfield', 'trans_tol': 1e-3},
    'config5_10k_frame': {'model': ('grid', 15, 15, 16, False), 'loads': 'gravity', 'subsets': 'connected', 'trans_tol': 1e-3},
    'config2_cantilever_bfs_prefixes': {'model': ('json', 'cantilever_beam_truss.json'), 'loads': 'json:cantilever_beam_truss_loadcase.json',
                                        'subsets': 'bfs_prefixes_spread:64', 'trans_tol': 1e-3},
}


def _shapes_worker(args):
    """Reference cost of one geometry variant: a new Model (coordinates jittered like bench.py does), a new
    checker (element tables: numpy_stiffness.py:579-614), loads, one full-structure solve."""
    seeds, kind = args
    import contextlib
    import copy
    import io
    import warnings
    import numpy as np
    warnings.simplefilter('ignore')
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(7, 7, 8)
    base = np.array([n.point for n in model.nodes], dtype=float)
    lc = LoadCase(gravity_load=GravityLoad([0, 0, -1]))
    n_ok = 0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        for s in seeds:
            pts = syn.shape_variant(base, s)
            mg = copy.deepcopy(model)
            for n, pt in zip(mg.nodes, pts):
                n.point = [float(x) for x in pt]
            if kind == 'reference':
                from oracle import ref_bridge
                sc = ref_bridge.make_ref_checker(mg, lc, 1e-3)
                n_ok += 1 if sc.solve([]) else 0
            else:
                from oracle.frame_oracle import FrameOracle
                orc = FrameOracle(mg)
                orc.set_loads(lc)
                n_ok += 1 if orc.solve([]).success else 0
    return time.perf_counter() - t0, len(seeds), n_ok


def shapes_record(n_per_proc=3):
    n_procs = os.cpu_count() or 1
    kind = 'reference' if reference_available() else 'port'
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    seeds = list(range(n_per_proc * n_procs))
    with ctx.Pool(n_procs) as pool:
        pool.map(_shapes_worker, [([s], kind) for s in seeds[:n_procs]])            # warm-up
        t0 = time.perf_counter()
        pool.map(_shapes_worker, [(seeds[i::n_procs], kind) for i in range(n_procs)])
        wall = time.perf_counter() - t0
    return {'value': len(seeds) / wall, 'unit': 'shapes/s', 'cores': n_procs, 'kind': kind,
            'sample': '{} shape variants of the 7x7x8 frame on {} processes, {:.1f} s: new Model + new checker + solve([]) each'.format(
                len(seeds), n_procs, wall)}


def other_configs_record():
    """CPU rates of configs 1, 2, 4 and 5 on bounded samples (a few seconds each)."""
    n_procs = os.cpu_count() or 1
    kind = 'reference' if reference_available() else 'port'
    out = {}
    out['config1_tower_full_solve'] = _rate_record(OTHER_CONFIGS['config1_tower_full_solve'], 50, 1, kind)
    out['config2_tower_bfs_prefixes'] = _rate_record(OTHER_CONFIGS['config2_tower_bfs_prefixes'], 24, 1, kind)
    out['config4_joints_4_loadcases'] = _rate_record(OTHER_CONFIGS['config4_joints_4_loadcases'], 6 * n_procs, n_procs, kind)
    out['config5_10k_frame'] = _rate_record(OTHER_CONFIGS['config5_10k_frame'], n_procs, n_procs, kind)
    out['config2_cantilever_bfs_prefixes'] = _rate_record(OTHER_CONFIGS['config2_cantilever_bfs_prefixes'], 64, n_procs, kind)
    out['config2_cantilever_bfs_prefixes']['sample'] += ' (64 prefixes evenly spread over the 2197-step sequence: the mean prefix)'
    out['many_shapes_one_topology'] = shapes_record()
    return out


if __name__ == '__main__':
    # python -m oracle.cpu_bench nx ny nz trans_tol batch sample [others] [parity=PATH,N]  ->  one JSON line (used by bench.py
    # in a child process so that the CPU run shares no state - threads, allocator - with the GPU process)
    import json
    import sys
    a = sys.argv[1:]
    rec = baseline_record((int(a[0]), int(a[1]), int(a[2])), float(a[3]), int(a[4]), int(a[5]))
    if 'others' in a[6:]:
        rec['other_configs'] = other_configs_record()
    for tok in a[6:]:
        if tok.startswith('parity='):                  # parity=PATH,N : the in-run parity set (never fails the baseline)
            path, n = tok[7:].rsplit(',', 1)
            try:
                spec = dict(CONFIG3, model=('grid', int(a[0]), int(a[1]), int(a[2]), False), trans_tol=float(a[3]))
                rec['parity_file'] = {'path': path, 'subsets': int(n), 'kind': parity_sample(spec, int(n), path)}
            except Exception as ex:
                rec['parity_file'] = {'error': repr(ex)}
    print(json.dumps(rec))
