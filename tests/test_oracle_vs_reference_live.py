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
from conmech_b200.io_base import (Node, Element, Support, Joint, CrossSec, Material, PointLoad, UniformlyDistLoad)
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


# ---- random small frames: everything the model format can express, mixed
def random_frame(rng):
    """4-12 nodes (half of the models snapped to a unit grid: axis-aligned and vertical members), random
    members with random orientation, tags and truss flags, full or partial fixities, joint releases / springs per tag
    (Q2), one or two materials and cross sections by tag (Q12)."""
    nV = rng.randint(4, 12)
    pts = [[rng.uniform(0, 3) for _ in range(3)] for _ in range(nV)]
    if rng.random() < 0.5:
        pts = sorted(set(tuple(float(round(x)) for x in p) for p in pts))
        pts = [list(p) for p in pts]
        rng.shuffle(pts)
        nV = len(pts)
        if nV < 3:
            return None
    nodes = [Node(p, i, False) for i, p in enumerate(pts)]
    pairs = [(i, j) for i in range(nV) for j in range(i + 1, nV)]
    rng.shuffle(pairs)
    nE = rng.randint(nV - 1, min(len(pairs), 3 * nV))
    elements = []
    for k, (i, j) in enumerate(pairs[:nE]):
        if rng.random() < 0.5:
            i, j = j, i
        elements.append(Element([i, j], k, rng.choice(['', 'a', 'b', 'c']), rng.random() > 0.2))
    supports = [Support([1] * 6 if rng.random() < 0.6 else [rng.randint(0, 1) for _ in range(6)], v)
                for v in rng.sample(range(nV), rng.randint(1, max(1, nV // 3)))]

    def spring():
        r = rng.random()
        return None if r < 0.5 else (0 if r < 0.8 else rng.choice([50.0, 1e3, 1e16]))
    joints = [Joint([None] * 3 + [spring(), spring(), spring()] + [None] * 3 + [spring(), spring(), spring()], [t])
              for t in ('a', 'b') if rng.random() < 0.7]
    mats = [Material(2.1e8, 8.076e7, 235e3, 78.5, elem_tags=None, family='S', name='S')]
    if rng.random() < 0.5:
        mats.append(Material(1.0e7, 4e6, 1e3, 5.0, elem_tags=['c'], family='W', name='W'))
    secs = [CrossSec(1e-3, 2e-7, 1e-7, 1.5e-7, elem_tags=None, family='r', name='r')]
    if rng.random() < 0.5:
        secs.append(CrossSec(2e-3, 4e-7, 3e-7, 2e-7, elem_tags=['b'], family='r', name='r2'))
    return Model(nodes, elements, supports, joints, mats, secs, model_name='fuzz')


def random_loadcase(rng, model):
    pls = [PointLoad([rng.gauss(0, 1) for _ in range(3)], [rng.gauss(0, 0.1) for _ in range(3)], v)
           for v in range(len(model.nodes)) if rng.random() < 0.3]
    uls = [UniformlyDistLoad([0, 0, -rng.random()], [0, 0, -0.1], [rng.choice(['a', 'b', 'c'])])] if rng.random() < 0.4 else []
    g = GravityLoad([0, 0, -1]) if (rng.random() < 0.6 or not (pls or uls)) else None
    return LoadCase(pls, uls, g)


def test_random_small_frames():
    """300 random frames x 5 subsets each (the full structure and four random element subsets - connected or not,
    mechanisms included): the oracle returns the reference's flag, has_result and DOF maps on every one, and its values
    on every well-posed one (SURVEY Q16 class W: min scaled pivot >= 1e-8)."""
    import random
    import warnings
    n_models = n_solves = n_values = n_true = n_noresult = n_rejected = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for seed in range(300):
            rng = random.Random(seed)
            model = random_frame(rng)
            if model is None:
                continue
            lc = random_loadcase(rng, model)
            nV, nE = len(model.nodes), len(model.elements)
            try:
                sc = ref_bridge.make_ref_checker(model, lc, 1e-2)
            except Exception:
                # a programmer error for the reference (e.g. a load on a tag no element carries: KeyError,
                # numpy_stiffness.py:117) is one for the oracle as well
                with pytest.raises(Exception):
                    FrameOracle(model).set_loads(lc)
                n_rejected += 1
                continue
            orc = FrameOracle(model)
            orc.set_loads(lc)
            orc.set_nodal_displacement_tol(1e-2)
            n_models += 1
            for ids in [[]] + [sorted(rng.sample(range(nE), rng.randint(1, nE))) for _ in range(4)]:
                r = reference_solve(sc, list(ids), nV, nE)
                o = orc.solve(list(ids))
                n_solves += 1
                assert o.success == r['flag'] and o.has_result == r['has'], (seed, ids)
                n_true += int(r['flag'])
                n_noresult += int(not r['has'])
                if 'dof_map' in r and o.dof_map is not None:
                    assert (o.n_free, o.n_fixed) == (r['n_free'], r['n_fixed']), (seed, ids)
                    assert_array_equal(o.dof_map, r['dof_map'])
                    assert_array_equal(o.node_exists, r['node_exists'])
                well_posed = np.isnan(o.min_pivot) or abs(o.min_pivot) >= 1e-8
                if r['has'] and well_posed and np.isfinite(r['U']).all() and np.abs(r['U']).max() < 1e6:
                    e_ids = list(ids) or list(range(nE))
                    assert split_err(o.U, r['U']) <= 1e-9 and split_err(o.R, r['R']) <= 1e-9, (seed, ids)
                    assert split_err(o.eR[e_ids], r['eR'][e_ids]) <= 1e-9, (seed, ids)
                    n_values += 1
    # the generator has to reach every outcome, or the test pins nothing
    assert n_models >= 250 and n_values >= 200 and n_true >= 50 and n_noresult >= 50, (n_models, n_rejected, n_solves, n_values, n_true, n_noresult)
