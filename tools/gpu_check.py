"""Diagnostic run on a GPU box: parity of the CUDA engine against the oracle on the golden cases
(both factorisation variants), then first timings.  Prints, does not assert."""
import os
import sys
import time
import traceback

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from golden_utils import Golden, golden_names, split_err, rel_err          # noqa: E402
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad              # noqa: E402
from conmech_b200 import synthetic as syn                                     # noqa: E402
from oracle.frame_oracle import FrameOracle                                   # noqa: E402

ALL = ('flag', 'status', 'n_free', 'n_fixed', 'dof_map', 'node_exists', 'U', 'Rf', 'eR', 'compliance',
       'max_trans', 'max_rot', 'min_pivot')


def compare(name, variants=(1, 6, 0)):
    g = Golden(name)
    model = g.model()
    lc = g.loadcase(model)
    orc = FrameOracle(model)
    orc.set_loads(lc)
    orc.set_nodal_displacement_tol(g.meta['trans_tol'])
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_loads(lc)
    sc.set_nodal_displacement_tol(g.meta['trans_tol'])
    eng = sc._sc_ins
    K = np.asarray(eng.get_element_stiffness_matrices())
    eK = rel_err(K, orc.K_G)
    subsets = [list(g.rec(i)['ids']) for i in range(g.n)]
    line = '{:45s} info={} Kg_err={:.1e}'.format(name, {k: v for k, v in eng.info.items() if k in ('n_free_full', 'half_bandwidth', 'band_tiles', 'n_tile_rows_max')}, eK)
    print(line)
    for var in variants:
        eng.set_factor_variant(var)
        r = eng.solve_many(subsets, want=ALL)
        worst = dict(U=0.0, R=0.0, eR=0.0, mt=0.0, c=0.0)
        bad = []
        for b, ids in enumerate(subsets):
            o = orc.solve(ids)
            ids2 = ids or list(range(orc.nE))
            f = bool(r['flag'][b, 0])
            st = int(r['status'][b, 0])
            if f != o.success:
                bad.append(('flag', b, f, o.success, st, o.status))
            if o.dof_map is not None:
                if int(r['n_free'][b]) != o.n_free or int(r['n_fixed'][b]) != o.n_fixed:
                    bad.append(('nfree', b, int(r['n_free'][b]), o.n_free))
                if not (r['dof_map'][b] == o.dof_map).all():
                    bad.append(('dofmap', b))
                if not (r['node_exists'][b].astype(bool) == o.node_exists).all():
                    bad.append(('exists', b))
            wellposed = o.has_result and np.isfinite(o.U).all() and np.abs(o.U).max() < 1e6
            if wellposed:
                if st not in (0, 4):
                    bad.append(('status', b, st, o.status, float(r['min_pivot'][b])))
                    continue
                worst['U'] = max(worst['U'], split_err(r['U'][b, 0], o.U))
                worst['R'] = max(worst['R'], split_err(r['Rf'][b, 0], o.R))
                worst['eR'] = max(worst['eR'], split_err(r['eR'][b, 0][ids2], o.eR[ids2]))
                worst['mt'] = max(worst['mt'], abs(r['max_trans'][b, 0] - o.max_trans) / max(o.max_trans, 1e-300))
                worst['c'] = max(worst['c'], abs(r['compliance'][b, 0] - o.compliance) / max(abs(o.compliance), 1e-300))
        print('    variant {}: '.format(var) + ' '.join('{}={:.1e}'.format(k, v) for k, v in worst.items()) +
              ('  BAD: ' + str(bad[:6]) if bad else '  ok'))


def timing():
    import torch
    model = syn.grid_frame(7, 7, 8)
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    eng = sc._sc_ins
    print('G1k info', eng.info)
    t0 = time.time()
    subsets = [syn.connected_subset(model, s) for s in range(4096)]
    print('sampler: {:.2f} ms/subset'.format((time.time() - t0) / 4096 * 1e3))
    masks = syn.pack_masks(subsets, len(model.elements))
    eng.enable_timing(True)
    for var, nb in ((2, 4096), (3, 4096)):
        eng.set_factor_variant(var)
        for rep in range(2):
            t0 = time.time()
            r = eng.solve_many(masks[:nb], want=('flag', 'status', 'max_trans'))
            dt = time.time() - t0
        print('variant {} B={}: wall {:.1f} ms -> {:.0f} solves/s; stage ms (assemble, factor, solve) = {}; flags true {}/{} status {}'.format(
            var, nb, dt * 1e3, nb / dt, eng.last_stage_ms(), int(r['flag'].sum()), nb, np.bincount(r['status'].ravel())))
    a = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
    b = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print('cuBLAS DGEMM 8192^3: {:.2f} ms -> {:.1f} TFLOP/s'.format(best, 2 * 8192**3 / best / 1e9))


if __name__ == '__main__':
    names = sys.argv[1:] or golden_names()
    for n in names:
        try:
            compare(n)
        except Exception:
            traceback.print_exc()
    try:
        timing()
    except Exception:
        traceback.print_exc()
