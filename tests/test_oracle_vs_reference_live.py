"""The oracle against the LIVE reference at the benchmark configurations themselves (CPU, no GPU).

tests/test_oracle_golden.py pins the oracle to vectors frozen from the reference; the frozen cases are small
where the configurations are large (joints on a 5x5x5 grid, not on the 7x7x8 frame of config 4).  Here the
unmodified reference from ``baseline/_ref`` (installed by ``__graft_entry__.build()`` when /root/reference is
present, shipped to the GPU box by gpurun) solves samples of configs 2-5 of BASELINE.json as SURVEY section 8d
defines them, and the oracle must agree: flags, has_result, n_free / n_fixed, DOF map and node set exactly,
values to 1e-12 (same scipy calls in the same order).  Skipped when the reference is not installed."""
import contextlib
import io

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from conmech_b200 import LoadCase, GravityLoad, Model
from conmech_b200 import synthetic as syn
from oracle import ref_bridge
from oracle.frame_oracle import FrameOracle
from golden_utils import MODELS, split_err, rel_err

pytestmark = pytest.mark.skipif(not ref_bridge.available(), reason='reference not installed under baseline/_ref')
TIGHT = 1e-12


def reference_solve(sc, ids, nV, nE):
    """One solve of the reference facade, results scattered to dense arrays by node / element id."""
    eng = sc._sc_ins
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            perm = eng.compute_dof_permutation(ids)
        except ValueError:
            perm = None
        try:
            flag = bool(sc.solve(ids))
        except TypeError:                      # bare False from compute_dof_permutation
            flag = False
            eng._has_stored_deformation = False
    rec = {'flag': flag, 'has': bool(sc.has_stored_result())}
    if isinstance(perm, tuple):
        Perm, (node_set, n_free, n_fixed, _) = perm
        ex = np.zeros(nV, dtype=bool)
        ex[list(node_set)] = True
        rec.update(n_free=n_free, n_fixed=n_fixed, dof_map=Perm.tocsr().indices.astype(np.int32), node_exists=ex)
    if rec['has']:
        _, nD, fR, eR = eng.get_solved_results()
        U, R, E = np.zeros(6 * nV), np.zeros(6 * nV), np.full((nE, 12), np.nan)
        for row in nD:
            U[6 * int(row[0]):6 * int(row[0]) + 6] = row[1:]
        for row in fR:
            R[6 * int(row[0]):6 * int(row[0]) + 6] = row[1:]
        for row in eR:
            E[int(row[0])] = row[1:]
        rec.update(U=U, R=R, eR=E, max_trans=eng.get_max_nodal_deformation()[0], compliance=eng.get_compliance())
    return rec


def check(model, loadcase, subsets, trans_tol, expect_split=False):
    nV, nE = len(model.nodes), len(model.elements)
    sc = ref_bridge.make_ref_checker(model, loadcase, trans_tol)       # fresh checker per load case
    orc = FrameOracle(model)
    orc.set_loads(loadcase)
    orc.set_nodal_displacement_tol(trans_tol)
    flags = []
    for ids in subsets:
        ids = list(ids)
        r = reference_solve(sc, ids, nV, nE)
        o = orc.solve(ids)
        assert o.success == r['flag'] and o.has_result == r['has']
        flags.append(r['flag'])
        if 'dof_map' in r:
            assert (o.n_free, o.n_fixed) == (r['n_free'], r['n_fixed'])
            assert_array_equal(o.dof_map, r['dof_map'])
            assert_array_equal(o.node_exists, r['node_exists'])
        if r['has'] and np.isfinite(r['U']).all() and np.abs(r['U']).max() < 1e6:
            e_ids = ids or list(range(nE))
            assert split_err(o.U, r['U']) <= TIGHT and split_err(o.R, r['R']) <= TIGHT
            assert split_err(o.eR[e_ids], r['eR'][e_ids]) <= TIGHT
            assert rel_err(o.max_trans, r['max_trans']) <= TIGHT and rel_err(o.compliance, r['compliance']) <= TIGHT
    if expect_split:
        assert 0 < sum(flags) < len(flags), 'the tolerance was meant to split the flags'
    return flags


def test_config3_connected_subsets_of_the_1k_frame():
    model = syn.grid_frame(7, 7, 8)
    lc = LoadCase(gravity_load=GravityLoad([0, 0, -1]))
    subsets = [syn.connected_subset(model, s) for s in range(24)]
    assert not any(check(model, lc, subsets, 1e-3))                      # bench tolerance: every flag False
    check(model, lc, subsets, 2e-2, expect_split=True)                   # a tolerance inside the max_trans range


@pytest.mark.parametrize('lc_index', [0, 1, 2, 3])
def test_config4_joints_and_four_load_cases_on_the_full_frame(lc_index):
    model = syn.grid_frame(7, 7, 8, joints=True)
    lc = syn.config4_loadcases(model)[lc_index]
    subsets = [syn.heightfield_subset(model, 7, 7, 8, seed=s) for s in range(10)]
    check(model, lc, subsets, 1e-3)


def test_config2_prefixes_of_the_largest_bundled_model():
    import os
    model = Model.from_json(os.path.join(MODELS, 'cantilever_beam_truss.json'))
    lc = LoadCase.from_json(os.path.join(MODELS, 'cantilever_beam_truss_loadcase.json'))
    order = syn.bfs_sequence(model)
    check(model, lc, [order[:k] for k in (40, 700, 2197)], 1e-3)


def test_config5_one_subset_of_the_10k_frame():
    model = syn.grid_frame(15, 15, 16)
    lc = LoadCase(gravity_load=GravityLoad([0, 0, -1]))
    check(model, lc, [syn.connected_subset(model, 0, lo=0.2, hi=0.3)], 1e-3)
