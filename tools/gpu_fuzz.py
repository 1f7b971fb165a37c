"""CUDA path against the oracle on random small frames (the generator of tests/test_oracle_vs_reference_live.py:
partial fixities, releases / springs, truss members, several materials / sections, mixed loads; full structure and
random element subsets, mechanisms included).  A tool, not part of the test suite: written in the last session of
round 2 after the GPU budget was spent, NOT YET RUN on a GPU - run it first thing when one is available:

    gpurun -- 'timeout 300 python tools/gpu_fuzz.py 0 300'

Flags / has_result / DOF maps must be equal; values within 1e-9 on well-posed solves (SURVEY Q16 class W);
marginal solves (|min scaled pivot| < 1e-8) are counted, not asserted."""
import os
import random
import sys
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
warnings.simplefilter('ignore')
from conmech_b200 import StiffnessChecker                                   # noqa: E402
from oracle.frame_oracle import FrameOracle                                  # noqa: E402
from golden_utils import split_err                                           # noqa: E402
from test_oracle_vs_reference_live import random_frame, random_loadcase      # noqa: E402

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
stats = dict(models=0, rejected=0, solves=0, values=0, marginal=0, flag=0, result=0, maps=0, value=0)
for seed in range(seed0, seed0 + count):
    rng = random.Random(seed)
    model = random_frame(rng)
    if model is None:
        continue
    lc = random_loadcase(rng, model)
    try:
        orc = FrameOracle(model)
        orc.set_loads(lc)
    except Exception:
        stats['rejected'] += 1
        continue
    orc.set_nodal_displacement_tol(1e-2)
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_loads(lc)
    sc.set_nodal_displacement_tol(trans_tol=1e-2)
    nE = len(model.elements)
    subsets = [[]] + [sorted(rng.sample(range(nE), rng.randint(1, nE))) for _ in range(4)]
    r = sc._sc_ins.solve_many(subsets, want=('flag', 'status', 'U', 'Rf', 'eR', 'n_free', 'n_fixed', 'dof_map', 'node_exists'))
    stats['models'] += 1
    for b, ids in enumerate(subsets):
        o = orc.solve(list(ids))
        stats['solves'] += 1
        marginal = not (np.isnan(o.min_pivot) or abs(o.min_pivot) >= 1e-8)
        stats['marginal'] += int(marginal)
        has = int(r['status'][b].ravel()[0]) in (0, 4)
        if not marginal and bool(r['flag'][b, 0]) != o.success:
            stats['flag'] += 1
            print('seed', seed, 'ids', ids, 'FLAG cuda', bool(r['flag'][b, 0]), 'oracle', o.success, 'status', r['status'][b], o.status)
        if not marginal and has != o.has_result:
            stats['result'] += 1
            print('seed', seed, 'ids', ids, 'RESULT cuda status', r['status'][b], 'oracle', o.status)
        if o.dof_map is not None and o.status == 0:
            if int(r['n_free'][b]) != o.n_free or int(r['n_fixed'][b]) != o.n_fixed or not np.array_equal(r['dof_map'][b], o.dof_map) \
                    or not np.array_equal(r['node_exists'][b].astype(bool), o.node_exists):
                stats['maps'] += 1
                print('seed', seed, 'ids', ids, 'MAPS')
        if o.has_result and has and not marginal and np.isfinite(o.U).all() and np.abs(o.U).max() < 1e6:
            e_ids = list(ids) or list(range(nE))
            err = max(split_err(r['U'][b, 0], o.U), split_err(r['Rf'][b, 0], o.R), split_err(r['eR'][b, 0][e_ids], o.eR[e_ids]))
            stats['values'] += 1
            if err > 1e-9:
                stats['value'] += 1
                print('seed', seed, 'ids', ids, 'VALUE rel err', err, 'min pivot', o.min_pivot)
    sc._sc_ins.close()
print(stats)
sys.exit(1 if (stats['flag'] or stats['result'] or stats['maps'] or stats['value']) else 0)
