"""CPU timing of the oracle (test/bench infrastructure only): the reference algorithm restated
in oracle/frame_oracle.py, run on the host cores over a bounded sample of a workload.

Used by bench.py for the ``cpu_baseline`` object and for ``--impl reference``.  The reference
itself (a Python package under /root/reference) cannot travel to the GPU box, so the oracle
port stands in for it (kind = "port").  The port vectorises the reference's interpreted
triple loop in ``assemble_global_stiffness_matrix`` (numpy_stiffness_assembly.py:379-387) and its
per-DOF Python loops; everything else (scipy csc products, SuperLU with the same options) is
the same call sequence, so the port is FASTER than the reference it stands for - the reported
CPU baseline is conservative.  DESIGN.md records the ratio measured in the authoring container.
"""
import os
import time

_STATE = {}


def _init_worker(grid, trans_tol):
    import warnings
    warnings.simplefilter('ignore')
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    from oracle.frame_oracle import FrameOracle
    model = syn.grid_frame(*grid)
    orc = FrameOracle(model)
    orc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    orc.set_nodal_displacement_tol(trans_tol)
    _STATE['orc'] = orc
    _STATE['model'] = model
    _STATE['syn'] = syn


def _solve_seeds(seeds):
    orc, model, syn = _STATE['orc'], _STATE['model'], _STATE['syn']
    subsets = [syn.connected_subset(model, s) for s in seeds]      # not timed
    t0 = time.perf_counter()
    n_ok = 0
    for ids in subsets:
        n_ok += 1 if orc.solve(ids).success else 0
    return time.perf_counter() - t0, len(seeds), n_ok


def time_oracle(grid, seeds, n_procs=1, trans_tol=1e-3):
    """Solve the subsets of ``seeds`` on ``n_procs`` processes (one oracle per worker).
    Returns (solves_per_second, wall_seconds_of_the_solve_phase, n_true_flags)."""
    seeds = list(seeds)
    if n_procs <= 1:
        _init_worker(grid, trans_tol)
        dt, n, ok = _solve_seeds(seeds)
        return n / dt, dt, ok
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    chunks = [seeds[i::n_procs] for i in range(n_procs)]
    chunks = [c for c in chunks if c]
    with ctx.Pool(len(chunks), initializer=_init_worker, initargs=(grid, trans_tol)) as pool:
        pool.map(_solve_seeds, [[s] for s in seeds[:len(chunks)]])          # warm-up: imports, first solve
        t0 = time.perf_counter()
        res = pool.map(_solve_seeds, chunks)
        wall = time.perf_counter() - t0
    n = sum(r[1] for r in res)
    ok = sum(r[2] for r in res)
    return n / wall, wall, ok


def cpu_model_name():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def baseline_record(grid, trans_tol, batch, sample=0):
    """The ``cpu_baseline`` object of bench.py: the oracle on all host cores over a bounded sample."""
    n_procs = os.cpu_count() or 1
    n_sample = sample or min(batch, 400 * n_procs)          # ~10-15 s of CPU work
    rate, wall, _ = time_oracle(grid, range(n_sample), n_procs=n_procs, trans_tol=trans_tol)
    rate1, wall1, _ = time_oracle(grid, range(min(n_sample, 48)), n_procs=1, trans_tol=trans_tol)
    return {'value': rate, 'unit': 'solves/s', 'cores': n_procs, 'kind': 'port',
            'sample': 'first {} subsets of the same workload on {} processes ({:.1f} s); '
                      'single process: {:.1f} solves/s; {}'.format(n_sample, n_procs, wall, rate1, cpu_model_name()),
            'single_core_value': rate1}


if __name__ == '__main__':
    # python -m oracle.cpu_bench nx ny nz trans_tol batch sample  ->  one JSON line (used by bench.py in a
    # child process so that the CPU run shares no state - threads, allocator - with the GPU process)
    import json
    import sys
    a = sys.argv[1:]
    print(json.dumps(baseline_record((int(a[0]), int(a[1]), int(a[2])), float(a[3]), int(a[4]), int(a[5]))))
