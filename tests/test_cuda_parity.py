"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same inputs, and against the golden vectors frozen from the reference.

Bar (BASELINE.json north_star): DOF maps, subset masks and success flags bit-exact; displacements
and reactions within 1e-9 relative (inf-norm, translations/forces and rotations/moments
separately)."""
import random

import numpy as np
import pytest
from numpy.testing import assert_array_equal, assert_almost_equal

from golden_utils import Golden, golden_names, split_err, REL_TOL, MODELS
from oracle.frame_oracle import FrameOracle, ST_OK, ST_ZERO_LOAD, ST_SINGULAR, ST_NO_SUPPORT

pytestmark = pytest.mark.gpu

ALL = ('flag', 'status', 'n_free', 'n_fixed', 'dof_map', 'node_exists', 'U', 'Rf', 'eR', 'compliance',
       'max_trans', 'max_rot', 'min_pivot')
WELL_POSED_PIVOT = 1e-8      # class W = reference min scaled pivot >= 1e-8


def make_pair(model, loadcase, trans_tol=1e-3):
    from conmech_b200 import StiffnessChecker
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_loads(loadcase)
    sc.set_nodal_displacement_tol(trans_tol=trans_tol)
    orc = FrameOracle(model)
    orc.set_loads(loadcase)
    orc.set_nodal_displacement_tol(trans_tol=trans_tol)
    return sc, orc


CLASS_COUNTS = {}      # tag -> {'W': n, 'Z': n, 'M': n, ...}: SURVEY Q16 classes, printed at session end (conftest)


def _count(tag, cls):
    CLASS_COUNTS.setdefault(tag, {}).setdefault(cls, 0)
    CLASS_COUNTS[tag][cls] += 1


def check_against_oracle(eng, orc, subsets, r, lc=0, tag=''):
    """Class W (well posed) -> exact flags/ints and 1e-9 values; class Z (exactly singular in the
    reference) -> both False; class M (marginal) -> counted and reported, must not claim success when
    the reference fails.  Returns the number of class-M solves."""
    n_marginal = 0
    for b, ids in enumerate(subsets):
        o = orc.solve(ids)
        ids2 = list(ids) or list(range(orc.nE))
        st = int(r['status'][b, lc])
        where = (tag, b)
        if o.dof_map is not None:
            assert int(r['n_free'][b]) == o.n_free, where
            assert int(r['n_fixed'][b]) == o.n_fixed, where
            assert_array_equal(r['dof_map'][b], o.dof_map, err_msg=str(where))
        assert_array_equal(r['node_exists'][b].astype(bool), o.node_exists, err_msg=str(where))
        if o.status in (ST_NO_SUPPORT, 2):
            assert st == o.status and not r['flag'][b, lc], where
            _count(tag, 'no_support')
            continue
        if o.status == ST_ZERO_LOAD:
            assert st == ST_ZERO_LOAD and bool(r['flag'][b, lc]) == o.success, where
            assert not r['U'][b, lc].any()
            _count(tag, 'zero_load')
            continue
        well_posed = o.status == ST_OK and (o.n_free == 0 or o.min_pivot >= WELL_POSED_PIVOT)
        if not well_posed:
            # singular / marginal in the reference: the flag must not claim success when the
            # reference fails; values are rounding-dependent there
            if o.status == ST_SINGULAR:
                _count(tag, 'Z')
            else:
                _count(tag, 'M')
                n_marginal += 1
            if not o.success:
                assert not r['flag'][b, lc], where
            continue
        _count(tag, 'W')
        assert st == ST_OK, (where, st, float(r['min_pivot'][b]))
        assert bool(r['flag'][b, lc]) == o.success, (where, float(r['max_trans'][b, lc]), o.max_trans)
        assert split_err(r['U'][b, lc], o.U) <= REL_TOL, where
        assert split_err(r['Rf'][b, lc], o.R) <= REL_TOL, where
        assert split_err(r['eR'][b, lc][ids2], o.eR[ids2]) <= REL_TOL, where
        absent = np.setdiff1d(np.arange(orc.nE), ids2)
        assert not r['eR'][b, lc][absent].any(), where
        assert abs(r['max_trans'][b, lc] - o.max_trans) <= REL_TOL * o.max_trans, where
        assert abs(r['compliance'][b, lc] - o.compliance) <= REL_TOL * abs(o.compliance), where
    return n_marginal


@pytest.mark.parametrize('variant', [1, 2, 6, 7, 8, 9, 10, 11, 12, 13, 0],
                         ids=['auto', 'w4', 'w16', 'w2', 'w3', 'w1', 'ring1', 'ring2', 'ring2w8', 'tmem', 'scalar'])
@pytest.mark.parametrize('name', golden_names())
def test_golden_cases(name, variant):
    g = Golden(name)
    model = g.model()
    sc, orc = make_pair(model, g.loadcase(model), g.meta['trans_tol'])
    eng = sc._sc_ins
    eng.set_factor_variant(variant)
    subsets = [list(g.rec(i)['ids']) for i in range(g.n)]
    r = eng.solve_many(subsets, want=ALL)
    check_against_oracle(eng, orc, subsets, r, tag=name)
    # and directly against the vectors frozen from the real reference
    for b in range(g.n):
        rec = g.rec(b)
        assert bool(r['flag'][b, 0]) == bool(rec['flag']), (name, b)
        if 'dof_map' in rec:
            assert_array_equal(r['dof_map'][b], rec['dof_map'])
        if 'U' in rec and np.isfinite(rec['U']).all() and np.abs(rec['U']).max() < 1e3 and 'joints__freeform' not in name:
            assert split_err(r['U'][b, 0], rec['U']) <= REL_TOL, (name, b)
            assert split_err(r['Rf'][b, 0], rec['R']) <= REL_TOL, (name, b)


@pytest.mark.parametrize('name', ['tower_3D__self_weight', 'grid_5x5x5_joints__lc0', 'grid_4x3x4_diag__gravity',
                                  'cantilever_beam_truss__loads', 'uniform_load_analytical_model__loads'])
def test_element_tables(name):
    """K1 against the oracle's element math and the reference's tables (<=1e-13)."""
    g = Golden(name)
    model = g.model()
    sc, orc = make_pair(model, g.loadcase(model))
    eng = sc._sc_ins
    K = np.asarray(eng.get_element_stiffness_matrices())
    R = np.asarray(eng.get_element_local2global_rot_matrices())
    scale = np.abs(orc.K_G).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(K - orc.K_G) / scale) <= 1e-13
    assert np.max(np.abs(R - orc.R12)) <= 1e-15
    assert_array_equal(eng.get_element2dof_id_map(), orc.id_map)
    if g.extra('K_G') is not None:
        assert np.max(np.abs(K - g.extra('K_G')) / scale) <= 1e-13
    ids = list(range(orc.nE))
    P = eng.get_lumped_nodal_loads(ids)
    Po = orc.load_vector(ids)
    assert np.max(np.abs(P - Po)) <= 1e-13 * max(1.0, np.abs(Po).max())


def test_bfs_prefix_flags_config2():
    """BASELINE config 2: every prefix of the construction sequence, tol 1e-3."""
    for name, expect in (('tower_3D__bfs_prefixes', '000000000001110000000111'), ('topopt-100__bfs_prefixes', '1' * 132)):
        g = Golden(name)
        model = g.model()
        sc, _ = make_pair(model, g.loadcase(model), 1e-3)
        order = g.meta['order']
        flags = sc.solve_many([order[:k] for k in range(1, len(order) + 1)])
        assert ''.join('1' if f else '0' for f in flags) == expect
        for b in range(g.n):
            rec = g.rec(b)
            assert flags[b] == bool(rec['flag'])


def test_facade_solve_matches_reference_tests():
    """The reference's own consistency test (tests/python/test_stiffness_checker.py:17-93): named
    success / failure subsets, shuffled id order, dict results."""
    from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad
    import os
    path = os.path.join(MODELS, 'tower_3D.json')
    sc = StiffnessChecker.from_json(path, checker_engine='cuda')
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    assert not sc.has_stored_result()
    sc.set_nodal_displacement_tol(trans_tol=0.04)
    ok_ids = [1, 12, 5, 10, 4, 11, 13, 3, 2, 0, 7]
    for ids, expect in (([], True), (ok_ids, True), ([0, 4, 7, 8, 9], False)):
        success = sc.solve(ids, eid_sanity_check=True)
        assert success == expect
        assert sc.has_stored_result()
        s2, nD, fR, eR = sc.get_solved_results()
        assert s2 == success
        assert len(nD) == len(sc.get_element_connected_node_ids(existing_ids=ids))
        assert len(eR) == (len(ids) or len(sc.elements))
        for fn in sc.get_element_connected_node_ids(existing_ids=ids, fix_node_only=True):
            assert_array_equal(nD[fn], np.zeros(6))
        if expect:
            assert sc.get_compliance() > 0.0
        rng = random.Random(0)
        for _ in range(5):
            sh = list(ids)
            rng.shuffle(sh)
            assert sc.solve(sh) == success
            _, nD2, fR2, eR2 = sc.get_solved_results()
            for i in nD:
                assert_almost_equal(nD[i], nD2[i])
            for i in fR:
                assert_almost_equal(fR[i], fR2[i])
            for i in eR:
                assert_almost_equal(eR[i][0], eR2[i][0])
                assert_almost_equal(eR[i][1], eR2[i][1])
    # no support in the subset -> False, no stored result (numpy_stiffness.py:238-240)
    assert sc.solve([4, 5, 6, 7]) is False
    assert not sc.has_stored_result()
    assert_array_equal(sc.fix_element_ids, [0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15])
    with pytest.raises(ValueError):
        sc._sc_ins.compute_dof_permutation([4, 5, 6, 7])
    with pytest.raises(NotImplementedError):
        StiffnessChecker.from_json(path, checker_engine='numpy')


def test_equilibrium_invariants():
    """Physics invariants the reference tests (test_stiffness_checker.py:191-365, 428-492):
    K u = f + r, nodal equilibrium of element end forces, sum of vertical reactions = weight."""
    g = Golden('topopt-100__self_weight+point_load')
    model = g.model()
    sc, orc = make_pair(model, g.loadcase(model))
    eng = sc._sc_ins
    ids = list(g.rec(1)['ids'])
    r = eng.solve_many([ids], want=ALL)
    U, R, E = r['U'][0, 0], r['Rf'][0, 0], r['eR'][0, 0]
    K = eng.compute_full_stiffness_matrix(ids)
    P = eng.get_lumped_nodal_loads(ids)
    ex = np.repeat(r['node_exists'][0].astype(bool), 6)
    assert np.max(np.abs(K.dot(U) - (P + R))[ex]) <= 1e-9 * np.abs(P).max()
    # fR + load = - sum_e R^T eR  at every existing node
    Rl = eng.get_element_local2global_rot_matrices()
    acc = np.zeros_like(U)
    for e in ids:
        acc[eng.get_element2dof_id_map()[e]] += Rl[e].T.dot(E[e])
    assert np.max(np.abs((R + P + acc)[ex])) <= 1e-8 * np.abs(P).max()
    # gravity only: vertical reactions carry the weight
    from conmech_b200 import LoadCase, GravityLoad
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    r = eng.solve_many([[]], want=('Rf', 'compliance'))
    W = eng.get_gravity_nodal_loads([]).reshape(-1, 6)[:, 2].sum()
    assert abs(r['Rf'][0, 0].reshape(-1, 6)[:, 2].sum() + W) <= 1e-9 * abs(W)
    c1 = r['compliance'][0, 0]
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -2])))
    c2 = eng.solve_many([[]], want=('compliance',))['compliance'][0, 0]
    assert abs(c2 / c1 - 4.0) <= 1e-9            # compliance scales with g^2


def test_edge_cases():
    from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, PointLoad, pack_masks
    g = Golden('tower_3D__self_weight')
    model = g.model()
    sc, orc = make_pair(model, LoadCase())          # no load at all
    eng = sc._sc_ins
    r = eng.solve_many([[], [0], [4, 5]], want=ALL)
    assert list(r['status'][:, 0]) == [ST_ZERO_LOAD, ST_ZERO_LOAD, ST_NO_SUPPORT]
    assert list(r['flag'][:, 0]) == [1, 1, 0]        # 0 < tol (numpy_stiffness.py:360-363)
    assert not r['U'][0].any() and not r['Rf'][0].any()
    # a load on a node that is absent from the subset still counts in ||P|| (Appendix B)
    sc.set_loads(LoadCase(point_loads=[PointLoad([0, 0, -1.0], [0, 0, 0], 9)]))
    orc.set_loads(LoadCase(point_loads=[PointLoad([0, 0, -1.0], [0, 0, 0], 9)]))
    subsets = [[0], [0, 1, 2, 3], []]
    r = eng.solve_many(subsets, want=ALL)
    check_against_oracle(eng, orc, subsets, r, tag='absent-load')
    assert r['status'][0, 0] == ST_OK and not r['U'][0, 0].any()
    # masks given directly; out-of-range ids raise; a repeated id raises (the reference would double count
    # the element, numpy_stiffness_assembly.py:371-388: a documented divergence)
    m = pack_masks([[0, 1, 2, 3]], len(model.elements))
    r2 = eng.solve_many(m, want=('U',))
    assert_array_equal(r2['U'][0], r['U'][1])
    r2 = eng.solve_many(m.view(np.int32), want=('U',))           # int32 view of the same masks (torch's dtype)
    assert_array_equal(r2['U'][0], r['U'][1])
    with pytest.raises(IndexError):
        eng.solve_many([[99]])
    with pytest.raises(ValueError):
        eng.solve_many([[0, 0, 1, 2, 3]])
    with pytest.raises(ValueError):
        sc.solve([2, 1, 2])
    assert eng.solve_many((ids for ids in subsets), want=('flag',))['flag'].shape == (3, 1)     # generators are fine
    # batch larger than one wave of the internal workspace still lines up
    big = [subsets[b % 3] for b in range(300)]
    rb = eng.solve_many(big, want=('max_trans',))
    for b in range(300):
        a, c = rb['max_trans'][b, 0], r['max_trans'][b % 3, 0]
        assert a == c or (np.isnan(a) and np.isnan(c))


def test_mechanism_detection():
    """Exactly singular systems (zero diagonal after assembly, class Z of SURVEY Q16): pin-ended
    members only -> rotational DOFs of free nodes have no stiffness; reference raises in SuperLU."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(3, 3, 3)
    for e in model.elements:
        e.bending_stiff = False
    sc, orc = make_pair(model, LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    r = sc._sc_ins.solve_many([[]], want=ALL)
    o = orc.solve([])
    assert o.status == ST_SINGULAR and not o.success
    assert r['status'][0, 0] == ST_SINGULAR and not r['flag'][0, 0]
    assert sc.solve([]) is False and not sc.has_stored_result()


def test_config3_grid_batch():
    """BASELINE config 3 shape: 1k-element grid, random connected subsets, vs the oracle; plus
    run-to-run determinism (no atomics on floating-point data)."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(7, 7, 8)
    sc, orc = make_pair(model, LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    eng = sc._sc_ins
    subsets = [syn.connected_subset(model, s) for s in range(1000, 1024)]
    r = eng.solve_many(subsets, want=ALL)
    check_against_oracle(eng, orc, subsets, r, tag='g1k')
    r2 = eng.solve_many(subsets, want=ALL)
    for k in ALL:
        assert_array_equal(r[k], r2[k])


def test_compact_results_match_dense():
    """Compact results (reactions of the support nodes, end forces of the existing elements only - the rows the
    reference's get_solved_results returns, numpy_stiffness.py:375-391) are bit for bit the dense arrays without their
    zero padding: host path over several waves, device path, several load cases, failing subsets, wrong offsets."""
    import torch
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad, PointLoad, _capi
    for model, lcs, B in ((syn.grid_frame(7, 7, 8), [LoadCase(gravity_load=GravityLoad([0, 0, -1]))], 3000),
                          (syn.grid_frame(4, 3, 4), [LoadCase(gravity_load=GravityLoad([0, 0, -1])),
                                                     LoadCase(point_loads=[PointLoad([1, 0, 0], [0, 0, 0], 40)]),
                                                     LoadCase(gravity_load=GravityLoad([0, 0, -1]), point_loads=[PointLoad([0, 2, 0], [0, 1, 0], 30)])], 200)):
        from conmech_b200 import StiffnessChecker
        sc = StiffnessChecker(model, checker_engine='cuda')
        eng = sc._sc_ins
        if len(lcs) == 1:
            sc.set_loads(lcs[0])
        else:
            sc.set_load_cases(lcs)
        nE, L = len(model.elements), len(lcs)
        rng = random.Random(5)
        subsets = [syn.connected_subset(model, s) for s in range(B - 3)]
        floating = [e for e in range(nE) if not (set(model.elements[e].end_node_inds) & set(model.supports.keys()))]
        subsets += [[], floating[:3], sorted(rng.sample(range(nE), nE // 2))]      # full structure, no support, arbitrary
        masks = syn.pack_masks(subsets, nE)
        dense = eng.solve_many(masks, want=('flag', 'status', 'U', 'Rf', 'eR'))
        comp = eng.solve_many(masks, want=('flag', 'status', 'U', 'Rf_sup', 'eR_rows'))
        off = comp['eR_row_offsets']
        assert_array_equal(off, eng.element_row_offsets(masks))
        assert comp['eR_rows'].shape == (off[-1] * L, 12)
        assert_array_equal(comp['flag'], dense['flag'])
        assert_array_equal(comp['U'], dense['U'])
        assert (dense['status'] != ST_OK).any() and (dense['status'] == ST_OK).any()
        sup = np.array(list(model.supports.keys()))
        assert_array_equal(comp['Rf_sup'], dense['Rf'].reshape(B, L, -1, 6)[:, :, sup])
        for b in range(B):
            ids = sorted(subsets[b]) or list(range(nE))
            rows = comp['eR_rows'][off[b] * L:off[b + 1] * L].reshape(L, len(ids), 12)
            assert_array_equal(rows, dense['eR'][b][:, ids], err_msg=str(b))
        # the device entry point delivers the same rows
        dm = torch.from_numpy(masks.view(np.int32)).cuda()
        rd = eng.solve_many_device(dm, want=('flag', 'Rf_sup', 'eR_rows'))
        torch.cuda.synchronize()
        assert_array_equal(rd['eR_row_offsets'].cpu().numpy(), off)
        assert_array_equal(rd['eR_rows'].cpu().numpy(), comp['eR_rows'])
        assert_array_equal(rd['Rf_sup'].cpu().numpy(), comp['Rf_sup'])
        # offsets that do not match the masks are refused before anything is written
        bad = off.copy()
        bad[B // 2] += 1
        oh = _capi.SolveOutHost()
        rows = np.zeros_like(comp['eR_rows'])
        oh.eR_rows, oh.eR_row_offsets = rows.ctypes.data, bad.ctypes.data
        import ctypes as C
        assert eng._lib.cm_solve_many_host(eng._handle, masks.ctypes.data, B, C.byref(oh)) != _capi.CM_OK
        assert b'eR_row_offsets' in eng._lib.cm_last_error()
        oh.eR_row_offsets = None
        assert eng._lib.cm_solve_many_host(eng._handle, masks.ctypes.data, B, C.byref(oh)) != _capi.CM_OK


def test_config3_full_size_properties():
    """Full-size batch (4096 subsets of the 1k frame): properties that need no oracle -
    vertical reactions balance the self weight of the existing elements, displacement at fixed
    DOFs is zero, flags equal (max_trans < tol), device path == host path."""
    import torch
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(7, 7, 8)
    sc, orc = make_pair(model, LoadCase(gravity_load=GravityLoad([0, 0, -1])), trans_tol=5e-2)
    eng = sc._sc_ins
    B = 4096
    subsets = [syn.connected_subset(model, s) for s in range(B)]
    masks = syn.pack_masks(subsets, len(model.elements))
    r = eng.solve_many(masks, want=('flag', 'status', 'U', 'Rf', 'max_trans', 'n_free'))
    assert (r['status'] == ST_OK).all()
    q = eng._element_loads(0)
    for b in range(0, B, 64):
        ids = subsets[b]
        w = q[ids][:, [2, 8]].sum()
        assert abs(r['Rf'][b, 0].reshape(-1, 6)[:, 2].sum() + w) <= 1e-9 * abs(w)
    sup = np.array(list(model.supports.keys()))
    assert not r['U'][:, 0].reshape(B, -1, 6)[:, sup].any()
    assert_array_equal(r['flag'][:, 0].astype(bool), r['max_trans'][:, 0] < 5e-2)
    assert 0 < r['flag'].sum() < B
    dm = torch.from_numpy(masks.view(np.int32)).cuda()
    rd = eng.solve_many_device(dm, want=('flag', 'max_trans', 'U'))
    torch.cuda.synchronize()
    assert_array_equal(rd['flag'].cpu().numpy(), r['flag'])
    assert_array_equal(rd['max_trans'].cpu().numpy(), r['max_trans'])
    assert_array_equal(rd['U'].cpu().numpy(), r['U'])


def test_solve_many_from_id_lists_pipelined():
    """The facade call with Python id lists (packing pipelined with the GPU calls chunk by chunk,
    one worker thread) returns exactly what the one-call path on pre-packed masks returns."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad, StiffnessChecker
    from conmech_b200.masks import pack_masks
    model = syn.grid_frame(4, 4, 4)
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    sc.set_nodal_displacement_tol(trans_tol=2e-5)
    eng = sc._sc_ins
    B = eng.PIPELINE_MIN + 777                      # above the threshold of the pipelined path
    lists = [syn.connected_subset(model, s) for s in range(B)]
    lists[5] = []
    want = ('flag', 'status', 'max_trans', 'compliance', 'n_free')
    r_lists = eng.solve_many(lists, want=want)
    r_masks = eng.solve_many(pack_masks(lists, len(model.elements)), want=want)
    for k in want:
        assert_array_equal(r_lists[k], r_masks[k])
    flags = sc.solve_many(lists)
    assert flags == [bool(f) for f in r_masks['flag'][:, 0]]
    flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in lists])
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])])
    assert sc.solve_many_csr(flat, off) == flags
    with pytest.raises(IndexError):
        eng.solve_many(lists[:B - 1] + [[len(model.elements)]], want=('flag',))
    assert eng.solve_many(lists[:3], want=('flag',))['flag'].shape == (3, 1)     # the handle is still usable


def _median_split_tol(values):
    """A tolerance in the widest gap near the median of `values`, so that `v < tol` splits them about
    50/50 and no value sits within rounding of the threshold."""
    v = np.sort(np.asarray(values, dtype=float))
    m = len(v) // 2
    lo, hi = max(1, m - len(v) // 10), min(len(v) - 1, m + len(v) // 10 + 1)
    k = lo + int(np.argmax(v[lo:hi] / v[lo - 1:hi - 1]))          # widest relative gap around the median
    tol = float(np.sqrt(v[k - 1] * v[k]))
    assert v[k - 1] * (1 + 1e-6) < tol < v[k] * (1 - 1e-6)
    return tol


def test_config4_full_frame_flag_split():
    """BASELINE config 4 as SURVEY 8d states it: the config-3 frame (7x7x8, 931 elements) with joint
    releases / springs on the horizontals, height-field subsets, FOUR load cases per subset solved in one
    pass, trans_tol at the oracle's median max_trans so that the success flags split about 50/50.
    100 subsets x 4 load cases against the oracle (fresh per load case, Q6): flags bit-exact, values 1e-9."""
    from conmech_b200 import synthetic as syn, StiffnessChecker
    dims = (7, 7, 8)
    model = syn.grid_frame(*dims, joints=True)
    lcs = syn.config4_loadcases(model)
    subsets = [syn.heightfield_subset(model, *dims, seed=s) for s in range(100)]
    orcs = []
    for lc in lcs:
        o = FrameOracle(model)
        o.set_loads(lc)
        orcs.append(o)
    mt = [o.solve(ids).max_trans for o in orcs for ids in subsets[:40]]
    tol = _median_split_tol([x for x in mt if np.isfinite(x) and x > 0])
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_load_cases(lcs)
    sc.set_nodal_displacement_tol(trans_tol=tol)
    eng = sc._sc_ins
    r = eng.solve_many(subsets, want=ALL)
    assert r['flag'].shape == (100, 4)
    n_m = 0
    for li, o in enumerate(orcs):
        o.set_nodal_displacement_tol(tol)
        n_m += check_against_oracle(eng, o, subsets, r, lc=li, tag='cfg4-7x7x8-lc{}'.format(li))
    assert n_m == 0                                    # the height-field sampler is class W by construction (Q16)
    frac = r['flag'].mean()
    assert 0.25 < frac < 0.75, frac                    # the flags do split


def test_config3_oracle_sample_flag_split():
    """Config 3 (1k-element frame, random connected subsets): 256 subsets against the oracle with
    trans_tol at the median max_trans of the sample - flag parity on a batch where the flags split."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(7, 7, 8)
    lc = LoadCase(gravity_load=GravityLoad([0, 0, -1]))
    subsets = [syn.connected_subset(model, s) for s in range(5000, 5256)]
    sc, orc = make_pair(model, lc)
    tol = _median_split_tol([orc.solve(ids).max_trans for ids in subsets[:64]])
    sc.set_nodal_displacement_tol(trans_tol=tol)
    orc.set_nodal_displacement_tol(tol)
    eng = sc._sc_ins
    r = eng.solve_many(subsets, want=ALL)
    n_m = check_against_oracle(eng, orc, subsets, r, tag='g1k-256')
    assert n_m == 0
    assert 0.25 < r['flag'].mean() < 0.75


def test_cantilever_bfs_prefixes():
    """Config 2 on the largest bundled model (cantilever_beam_truss: 817 nodes / 2197 elements, 68 point
    loads): 64 prefixes of the BFS construction sequence spread over its whole length, against the
    oracle; class-M prefixes (free-standing truss members early in the sequence) are counted."""
    import os
    from conmech_b200 import synthetic as syn, Model, LoadCase
    model = Model.from_json(os.path.join(MODELS, 'cantilever_beam_truss.json'))
    lc = LoadCase.from_json(os.path.join(MODELS, 'cantilever_beam_truss_loadcase.json'))
    sc, orc = make_pair(model, lc, 1e-3)
    order = syn.bfs_sequence(model)
    assert len(order) == len(model.elements)
    ks = sorted(set(int(round(x)) for x in np.linspace(1, len(order), 64)))
    subsets = [order[:k] for k in ks]
    eng = sc._sc_ins
    r = eng.solve_many(subsets, want=ALL)
    check_against_oracle(eng, orc, subsets, r, tag='cantilever-bfs')
    flags = sc.solve_sequence(order)
    assert [flags[k - 1] for k in ks] == [bool(f) for f in r['flag'][:, 0]]


def _compare_results(a, b, tag):
    """Two result sets of the same subsets from different solver orderings: integers / flags equal, values 1e-9."""
    for k in ('flag', 'n_free', 'n_fixed', 'node_exists'):
        assert_array_equal(a[k], b[k], err_msg='{} {}'.format(tag, k))
    st_a, st_b = a['status'][:, 0], b['status'][:, 0]
    assert_array_equal(st_a, st_b, err_msg=tag)
    has_map = ~np.isin(st_a, (ST_NO_SUPPORT, 2))           # no DOF map without supports (the reference raises there)
    assert_array_equal(a['dof_map'][has_map], b['dof_map'][has_map], err_msg=tag + ' dof_map')
    for i in np.flatnonzero((st_a == ST_OK) | (st_a == ST_ZERO_LOAD)):
        for k in ('U', 'Rf', 'eR'):
            assert split_err(a[k][i, 0].ravel(), b[k][i, 0].ravel()) <= REL_TOL, (tag, k, i)
        for k in ('max_trans', 'compliance'):
            assert abs(a[k][i, 0] - b[k][i, 0]) <= REL_TOL * max(abs(b[k][i, 0]), 1e-300), (tag, k, i)


def test_sequence_factor_reuse():
    """solve_sequence reuses the factorisation along a construction sequence (node ordering by first
    appearance, groups of prefixes sharing the factor rows of the group's first prefix).  Every prefix must come out as
    from the independent batched path (flags / status / integer outputs equal, values 1e-9) and as the oracle says - for
    BFS sequences, for a shuffled sequence (early prefixes without support: no sharing possible, fall-back paths), for
    several group sizes (1 = no sharing at all) and with joint releases (singular prefixes)."""
    import os
    from conmech_b200 import synthetic as syn, Model, LoadCase, GravityLoad, StiffnessChecker
    cases = []
    tower = Model.from_json(os.path.join(MODELS, 'tower_3D.json'))
    cases.append(('tower', tower, LoadCase(gravity_load=GravityLoad([0, 0, -1])), syn.bfs_sequence(tower)))
    shuffled = list(range(len(tower.elements)))
    random.Random(5).shuffle(shuffled)
    cases.append(('tower-shuffled', tower, LoadCase(gravity_load=GravityLoad([0, 0, -1])), shuffled))
    topo = Model.from_json(os.path.join(MODELS, 'topopt-100.json'))
    cases.append(('topopt', topo, Golden('topopt-100__self_weight+point_load').loadcase(topo), syn.bfs_sequence(topo)))
    grid = syn.grid_frame(4, 4, 4, joints=True, diagonals=True)
    cases.append(('grid-joints', grid, LoadCase(gravity_load=GravityLoad([0, 0, -1])), syn.bfs_sequence(grid)))
    for name, model, lc, order in cases:
        sc, orc = make_pair(model, lc, 1e-3)
        eng = sc._sc_ins
        prefixes = [order[:k] for k in range(1, len(order) + 1)]
        r_ind = eng.solve_many(prefixes, want=ALL)
        for group in (1, 5, 16):
            r_seq = eng.solve_sequence(order, want=ALL, group=group)
            _compare_results(r_seq, r_ind, '{}-g{}'.format(name, group))
        step = max(1, len(prefixes) // 24)
        pick = list(range(0, len(prefixes), step))
        check_against_oracle(eng, orc, [prefixes[i] for i in pick], {k: v[pick] for k, v in r_seq.items()}, tag='seq-' + name)
        sc.SEQUENCE_REUSE_MIN = 1                   # (the facade only reuses for long sequences; force it here)
        assert sc.solve_sequence(order) == [bool(f) for f in r_ind['flag'][:, 0]]
        assert sc.solve_sequence(order, reuse=False) == sc.solve_sequence(order)
    # the largest bundled model: all 2197 prefixes, reuse against the independent path
    model = Model.from_json(os.path.join(MODELS, 'cantilever_beam_truss.json'))
    lc = LoadCase.from_json(os.path.join(MODELS, 'cantilever_beam_truss_loadcase.json'))
    sc, orc = make_pair(model, lc, 1e-3)
    eng = sc._sc_ins
    order = syn.bfs_sequence(model)
    want = ('flag', 'status', 'n_free', 'n_fixed', 'max_trans', 'compliance')
    r_seq = eng.solve_sequence(order, want=want)
    r_ind = eng.solve_many([order[:k] for k in range(1, len(order) + 1)], want=want)
    for k in ('flag', 'status', 'n_free', 'n_fixed'):
        assert_array_equal(r_seq[k], r_ind[k], err_msg=k)
    ok = (r_ind['status'][:, 0] == ST_OK) & (r_ind['max_trans'][:, 0] > 0)
    assert ok.sum() > 400            # (point loads only: the early prefixes carry no load on their own nodes, U = 0)
    assert np.max(np.abs(r_seq['max_trans'][ok] - r_ind['max_trans'][ok]) / r_ind['max_trans'][ok]) <= REL_TOL
    ks = [10, 500, 1234, 2197]
    check_against_oracle(eng, orc, [order[:k] for k in ks],
                         {k: v[[x - 1 for x in ks]] for k, v in eng.solve_sequence(order, want=ALL).items()}, tag='seq-cantilever')


def test_k2_shared_memory_window():
    """a model whose K2 dynamic shared memory falls between 44.9 and 48 KB (static + dynamic
    above the 48 KB no-opt-in limit) must launch.  Grids around 2.5k nodes sit in that window."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    for dims in ((13, 14, 14), (13, 13, 15), (12, 14, 15)):
        model = syn.grid_frame(*dims)
        sc, orc = make_pair(model, LoadCase(gravity_load=GravityLoad([0, 0, -1])))
        eng = sc._sc_ins
        subsets = [syn.connected_subset(model, 1, lo=0.3, hi=0.4)]
        r = eng.solve_many(subsets, want=ALL)
        if dims == (13, 14, 14):
            check_against_oracle(eng, orc, subsets, r, tag='k2-smem-window')
        assert r['status'][0, 0] == ST_OK


def test_two_gpus_one_process_bit_exact():
    """One process driving two GPUs (devices=[0, 1]: cm_solve_many_host_multi, one host thread per
    device, results written straight into the caller's arrays) returns bit for bit what one GPU returns;
    so does the pipelined id-list path."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad, StiffnessChecker
    model = syn.grid_frame(7, 7, 8)
    lc = LoadCase(gravity_load=GravityLoad([0, 0, -1]))
    B = 6001
    lists = [syn.connected_subset(model, s) for s in range(B)]
    masks = syn.pack_masks(lists, len(model.elements))
    want = ('flag', 'status', 'max_trans', 'compliance', 'U', 'Rf', 'eR', 'n_free', 'Rf_sup', 'eR_rows')
    res = []
    for devices in ([0], [0, 1], [1]):
        sc = StiffnessChecker(model, checker_engine='cuda', devices=devices)
        sc.set_loads(lc)
        sc.set_nodal_displacement_tol(trans_tol=5e-2)
        res.append(sc._sc_ins.solve_many(masks, want=want))
        if len(devices) == 2:
            r_lists = sc._sc_ins.solve_many(lists, want=('flag', 'max_trans'))
            assert_array_equal(r_lists['flag'], res[-1]['flag'])
            assert_array_equal(r_lists['max_trans'], res[-1]['max_trans'])
            assert sc.solve(lists[7]) == bool(res[-1]['flag'][7, 0])
        sc._sc_ins.close()
    for k in want + ('eR_row_offsets',):
        assert_array_equal(res[0][k], res[1][k], err_msg=k)
        assert_array_equal(res[0][k], res[2][k], err_msg=k)
    off = res[1]['eR_row_offsets']           # the compact rows of both devices' shares are the dense rows of the existing elements
    for b in (0, B // 2 - 1, B // 2, B // 2 + 1, B - 1):
        assert_array_equal(res[1]['eR_rows'][off[b]:off[b + 1]], res[1]['eR'][b, 0][sorted(lists[b])], err_msg=str(b))


def test_geometry_variants_many_models():
    """many models of one topology.  G jittered shapes of a jointed grid (the element kernel
    builds all G x nE tables in one launch), a subset per shape, two load cases - against the oracle built
    from a fresh Model per shape, as a caller of the reference would do it."""
    import copy
    from conmech_b200 import synthetic as syn, StiffnessChecker, LoadCase, GravityLoad, PointLoad
    dims = (4, 4, 5)
    model = syn.grid_frame(*dims, joints=True, diagonals=True)
    nV, nE = len(model.nodes), len(model.elements)
    rng = np.random.RandomState(3)
    base = np.array([n.point for n in model.nodes], dtype=float)
    G = 7
    shapes = np.stack([base] + [base + rng.uniform(-0.12, 0.12, size=base.shape) for _ in range(G - 1)])
    lcs = [LoadCase(gravity_load=GravityLoad([0, 0, -1])),
           LoadCase(point_loads=[PointLoad([0.3, -0.2, -1.0], [0, 0, 0], nV - 1)], gravity_load=GravityLoad([0.1, 0, -1]))]
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_load_cases(lcs)
    eng = sc._sc_ins
    eng.set_geometries(shapes)
    assert eng.n_geometries() == G
    subsets = [[]] + [syn.heightfield_subset(model, *dims, seed=s) for s in range(40, 40 + G - 1)]
    geom = np.arange(G, dtype=np.int32)
    r = eng.solve_many(subsets, want=ALL, geometry=geom)
    # the same batch on device buffers, variants in reverse order
    import torch
    masks = syn.pack_masks(subsets, nE)[::-1].copy()
    rd = eng.solve_many_device(torch.from_numpy(masks.view(np.int32)).cuda(), want=('U', 'flag'),
                               geometry=torch.from_numpy(geom[::-1].copy()).cuda())
    torch.cuda.synchronize()
    assert_array_equal(rd['U'].cpu().numpy()[::-1], r['U'])
    n_checked = 0
    for g in range(G):
        mg = copy.deepcopy(model)
        for n, pt in zip(mg.nodes, shapes[g]):
            n.point = [float(x) for x in pt]
        for li, lc in enumerate(lcs):
            orc = FrameOracle(mg)
            orc.set_loads(lc)
            rg = {k: v[g:g + 1] for k, v in r.items()}
            check_against_oracle(eng, orc, [subsets[g]], rg, lc=li, tag='shapes-g{}-lc{}'.format(g, li))
            n_checked += 1
    assert n_checked == 2 * G
    # facade: one flag per shape; a collapsed member is a programmer error like in the reference (ValueError)
    sc2 = StiffnessChecker(model, checker_engine='cuda')
    sc2.set_loads(lcs[0])
    flags, det = sc2.solve_shapes(shapes, return_details=True)
    assert len(flags) == G and det['max_trans'][0] == r['max_trans'][0, 0]      # shape 0, full structure, load case 0
    assert_array_equal(det['status'], np.zeros(G, dtype=np.uint8))
    bad = shapes.copy()
    bad[2, model.elements[0].end_node_inds[0]] = bad[2, model.elements[0].end_node_inds[1]]
    with pytest.raises(ValueError):
        sc2.solve_shapes(bad)
    with pytest.raises(Exception):
        eng.solve_many([[]], geometry=[G])          # no such variant


@pytest.mark.parametrize('n_lc', [4, 3, 2])
def test_config4_multi_loadcase(n_lc):
    """BASELINE config 4: joints on horizontals, height-field subsets, several load cases solved
    in ONE pass against the oracle solved once per (subset, load case).  4 / 3 / 2 load cases
    exercise the back substitution with four right-hand sides per pass over L, a partly filled
    group of four, and two per pass."""
    from conmech_b200 import synthetic as syn, StiffnessChecker
    dims = (5, 5, 5)
    model = syn.grid_frame(*dims, joints=True)
    lcs = syn.config4_loadcases(model)[:n_lc]
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_load_cases(lcs)
    eng = sc._sc_ins
    subsets = [[]] + [syn.heightfield_subset(model, *dims, seed=s) for s in range(100, 112)]
    r = eng.solve_many(subsets, want=ALL)
    assert r['flag'].shape == (len(subsets), n_lc)
    for li, lc in enumerate(lcs):
        orc = FrameOracle(model)                    # fresh per load case (Q6)
        orc.set_loads(lc)
        n_m = check_against_oracle(eng, orc, subsets, r, lc=li, tag='cfg4-lc{}'.format(li))
        assert n_m == 0


def test_config5_large_frame():
    """BASELINE config 5: ~10k-element grid (15x15x16: 9675 elements, 20 250 free DOFs, band of 35
    tiles, 633 tile rows) - a factorisation-bound case; two subsets against the oracle."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(15, 15, 16)
    sc, orc = make_pair(model, LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    eng = sc._sc_ins
    assert eng.info['n_free_full'] == 20250
    subsets = [syn.connected_subset(model, 3, lo=0.5, hi=0.6), []]
    r = eng.solve_many(subsets, want=ALL)
    check_against_oracle(eng, orc, subsets, r, tag='g10k')


def test_solve_sequence_and_from_data():
    """Next-row callers: every prefix of a construction sequence in one call, and the
    one-call JSON-data entry point (reference utils.py:6-64) against the reference's stored result."""
    import json
    from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn
    from conmech_b200.utils import solve_linear_elastic_from_data
    g = Golden('tower_3D__bfs_prefixes')
    model = g.model()
    sc = StiffnessChecker(model, checker_engine='cuda')
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    sc.set_nodal_displacement_tol(trans_tol=1e-3)
    order = syn.bfs_sequence(model)
    flags = sc.solve_sequence(order)
    assert ''.join('1' if f else '0' for f in flags) == '000000000001110000000111'
    assert flags == sc.solve_many([order[:k] for k in range(1, len(order) + 1)])
    # cantilever_beam_truss: model + load case as data -> result dictionary like the reference's,
    # checked against the result file an older release of the reference stored with the example
    import os
    g2 = Golden('cantilever_beam_truss__loads')
    with open(os.path.join(MODELS, 'cantilever_beam_truss.json')) as f:
        mdata = json.load(f)
    with open(os.path.join(MODELS, 'cantilever_beam_truss_loadcase.json')) as f:
        ldata = json.load(f)
    rd = solve_linear_elastic_from_data(mdata, ldata, verbose=False)
    assert rd['solve_success'] is True
    assert abs(rd['elastic_energy'] - float(g2.extra('stored_compliance'))) <= 1e-9 * float(g2.extra('stored_compliance'))
    assert abs(rd['max_trans'] - float(g2.extra('stored_max_trans'))) <= 1e-9 * float(g2.extra('stored_max_trans'))
    U = g2.extra('stored_U').reshape(-1, 6)
    for nid, nd in rd['nodal_displacement'].items():
        assert np.max(np.abs(np.array(nd) - U[nid])) <= 1e-9 * np.abs(U).max()


def test_config3_100k_subsets_properties():
    """BASELINE config 3 at its FULL size - 100 000 random connected subsets of the 1k-element frame -
    through properties that need no oracle: every subset solves (status OK), the vertical support
    reactions balance the self weight of exactly the existing elements (1e-9 relative), flags equal
    (max_trans < tol), compliance is positive; plus the oracle on a sample of the same batch."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(7, 7, 8)
    tol = 5e-2
    sc, orc = make_pair(model, LoadCase(gravity_load=GravityLoad([0, 0, -1])), trans_tol=tol)
    eng = sc._sc_ins
    nE = len(model.elements)
    B, chunk = 100000, 5000
    q = eng._element_loads(0)
    wz = q[:, 2] + q[:, 8]                                    # vertical lumped self weight per element
    n_true = 0
    for c0 in range(0, B, chunk):
        subsets = [syn.connected_subset(model, s) for s in range(c0, c0 + chunk)]
        masks = syn.pack_masks(subsets, nE)
        r = eng.solve_many(masks, want=('flag', 'status', 'Rf', 'max_trans', 'compliance'))
        assert (r['status'] == ST_OK).all()
        bits = np.unpackbits(masks.view(np.uint8), axis=1, bitorder='little')[:, :nE].astype(np.float64)
        weight = bits @ wz
        rz = r['Rf'][:, 0].reshape(chunk, -1, 6)[:, :, 2].sum(axis=1)
        assert np.max(np.abs(rz + weight) / np.abs(weight)) <= 1e-9
        assert_array_equal(r['flag'][:, 0].astype(bool), r['max_trans'][:, 0] < tol)
        assert (r['compliance'] > 0).all()
        n_true += int(r['flag'].sum())
        if c0 == 0:
            sample = list(range(0, chunk, chunk // 16))
            rs = eng.solve_many([subsets[b] for b in sample], want=ALL)
            check_against_oracle(eng, orc, [subsets[b] for b in sample], rs, tag='g1k-100k-sample')
    assert 0 < n_true < B


def test_sigma_max_on_device():
    """the stress post-processing of the facade, computed in K4 for every (subset, load
    case): per-element sigma_max and its maximum magnitude, against the reference formula applied
    to the ORACLE's element reactions."""
    from conmech_b200 import synthetic as syn, LoadCase, GravityLoad
    model = syn.grid_frame(4, 4, 4)
    lc = LoadCase(gravity_load=GravityLoad([0, 0, -1]))
    sc, orc = make_pair(model, lc)
    eng = sc._sc_ins
    subsets = [[]] + [syn.connected_subset(model, s) for s in range(24)]
    r = eng.solve_many(subsets, want=('flag', 'status', 'eR', 'sigma_max', 'max_abs_sigma'))
    nE = len(model.elements)
    A = np.array([sc.get_element_crosssec(e).A for e in range(nE)])
    Iy = np.array([sc.get_element_crosssec(e).Iy for e in range(nE)])
    W = Iy / np.sqrt(A / np.pi)
    checked = 0
    for b, ids in enumerate(subsets):
        o = orc.solve(ids)
        if not o.has_result:
            continue
        ids2 = list(ids) or list(range(nE))
        eR = o.eR
        sigma = eR[:, 0] / A
        mag = np.abs(sigma) + np.maximum(np.abs(eR[:, 4]), np.abs(eR[:, 10])) / W + np.maximum(np.abs(eR[:, 5]), np.abs(eR[:, 11])) / W
        ref = np.where(sigma > 0, mag, -mag)
        got = r['sigma_max'][b, 0]
        scale = np.max(np.abs(ref[ids2]))
        # the sign is that of N/A: rounding noise for members without axial force, so magnitudes are
        # compared everywhere and signs where the axial stress is not negligible
        assert np.max(np.abs(np.abs(got[ids2]) - np.abs(ref[ids2]))) <= REL_TOL * scale, b
        firm = [e for e in ids2 if abs(sigma[e]) > 1e-6 * scale]
        assert np.array_equal(np.sign(got[firm]), np.sign(ref[firm])), b
        absent = np.setdiff1d(np.arange(nE), ids2)
        assert not got[absent].any()
        assert abs(r['max_abs_sigma'][b, 0] - np.max(np.abs(got))) == 0.0
        checked += 1
    assert checked >= 20
    # the facade's host-side accessor agrees with the device array for a single solve
    ids = subsets[3]
    sc.solve(ids)
    host = sc.get_sigma_max_per_element()
    for e, v in host.items():
        assert abs(abs(v) - abs(r['sigma_max'][3, 0][e])) <= REL_TOL * max(abs(v), 1e-300) + 1e-12 * scale


def test_c_client_cantilever_analytic(tmp_path):
    """A plain C host (tests/c_client/cantilever.c: no Python, no torch) drives the C ABI - cm_create, cm_set_loadcases,
    cm_element_row_offsets, cm_solve_many_host with dense and compact outputs - and checks the Euler-Bernoulli closed
    forms of a cantilever (tip deflection P L^3 / 3 E I, clamp reactions), the zero-load and no-support statuses."""
    import os
    import subprocess
    from test_host_cpu import _build_c_client, REPO
    exe = _build_c_client(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REPO, 'conmech_b200', 'lib'))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith('OK')
