"""Pin the CPU oracle to the reference: every golden vector frozen from the real
reference (tests/golden/make_golden.py), the reference's own stored result for its largest
example, and the textbook answers its tests hold."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal, assert_allclose

from oracle.frame_oracle import FrameOracle, ST_OK
from golden_utils import Golden, golden_names, rel_err, split_err

TIGHT = 1e-12     # same scipy calls in the same order: in practice bit-identical


def make_oracle(g):
    model = g.model()
    orc = FrameOracle(model)
    orc.set_loads(g.loadcase(model))
    orc.set_nodal_displacement_tol(trans_tol=g.meta['trans_tol'])
    return orc


@pytest.mark.parametrize('name', golden_names())
def test_oracle_matches_reference_golden(name):
    g = Golden(name)
    orc = make_oracle(g)
    K = g.extra('K_G')
    if K is not None:
        assert rel_err(orc.K_G, K) <= 1e-15
        assert_array_equal(orc.R3, g.extra('R3'))
        assert_array_equal(orc.id_map, g.extra('id_map'))
    if g.extra('q_gravity') is not None:
        assert_array_equal(orc.q_gravity, g.extra('q_gravity'))
    for i in range(g.n):
        r = g.rec(i)
        res = orc.solve(list(r['ids']))
        assert res.success == bool(r['flag']), (name, i)
        assert res.has_result == bool(r['has_result']), (name, i)
        if int(r['n_free']) >= 0:
            assert res.n_free == int(r['n_free']) and res.n_fixed == int(r['n_fixed'])
        if 'dof_map' in r:
            assert_array_equal(res.dof_map, r['dof_map'])
            assert_array_equal(res.node_exists, r['node_exists'])
        if 'max_trans' in r:
            assert_allclose(res.max_trans, float(r['max_trans']), rtol=TIGHT)
            assert_allclose(res.compliance, float(r['compliance']), rtol=TIGHT, atol=1e-300)
        if 'U' in r and np.isfinite(r['U']).all() and np.abs(r['U']).max() < 1e6:
            assert split_err(res.U, r['U']) <= TIGHT
            assert split_err(res.R, r['R']) <= TIGHT
            ids = list(r['ids']) or list(range(orc.nE))
            assert split_err(res.eR[ids], r['eR'][ids]) <= TIGHT
            assert np.isnan(res.eR[[e for e in range(orc.nE) if e not in set(ids)]]).all()


def test_stored_example_result():
    """examples/scripts/cantilever_beam_truss/*_result.json, written by an older release of the
    reference: displacements/reactions agree, element reactions are MINUS today's (CHANGELOG.rst:21)."""
    g = Golden('cantilever_beam_truss__loads')
    res = make_oracle(g).solve([])
    assert np.max(np.abs(res.U - g.extra('stored_U'))) < 1e-13
    assert np.max(np.abs(res.R - g.extra('stored_R'))) < 1e-10
    assert np.max(np.abs(res.eR + g.extra('stored_eR_presignflip'))) < 1e-9
    assert_allclose(res.compliance, float(g.extra('stored_compliance')), rtol=1e-12)
    assert_allclose(res.max_trans, float(g.extra('stored_max_trans')), rtol=1e-12)


def test_survey_known_answers():
    """SURVEY §8c golden (2): numbers measured on the reference during the survey."""
    res = make_oracle(Golden('tower_3D__self_weight')).solve([])
    assert_allclose(res.max_trans, 1.0717161129723429e-04, rtol=1e-12)
    assert_allclose(res.compliance, 8.939090469726433e-03, rtol=1e-12)
    res = make_oracle(Golden('topopt-100__self_weight')).solve([])
    assert_allclose(res.max_trans, 3.7790726973191865e-07, rtol=1e-12)
    assert_allclose(res.compliance, 2.4374697283631682e-11, rtol=1e-11)
    g = Golden('tower_3D__bfs_prefixes')
    assert g.meta['order'] == [0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15, 4, 7, 17, 22, 5, 19, 20, 6, 18, 21, 16, 23]
    orc = make_oracle(g)
    flags = ''.join('1' if orc.solve(g.meta['order'][:k]).success else '0' for k in range(1, 25))
    assert flags == '000000000001110000000111'


def test_textbook_mgz_5_8():
    """MGZ example 5.8 through the engine (reference tests/python/test_stiffness_checker.py:573-632)."""
    g = Golden('uniform_load_analytical_model__loads')
    orc = make_oracle(g)
    res = orc.solve([])
    U = res.U.reshape(-1, 6)
    R = res.R.reshape(-1, 6)
    assert_allclose(U[1], [0, 0, -0.02237, 4.195e-3, 5.931e-3, 0], atol=0.05)
    assert_allclose(R[0], [0, 0, 14.74, -6.45e-3, -36.21, 0], atol=0.05)
    assert_allclose(R[2], [0, 0, 5.25, -41.94, -17.11e-3, 0], atol=0.05)
    P = orc.load_vector(range(orc.nE)).reshape(-1, 6)
    assert P[1][2] == -12.5 and P[1][4] == -6.25


def test_textbook_mgz_4_8():
    """MGZ example 4.8 (reference tests/python/test_np_stiffness.py:61-112)."""
    res = make_oracle(Golden('mgz_flat_beam__loads')).solve([])
    U = res.U.reshape(-1, 6)
    R = res.R.reshape(-1, 6)
    assert res.status == ST_OK
    got = np.array([U[1][0], U[2][0], U[2][1], U[1][5], U[2][5]])
    assert_allclose(got, [0.024e-3, 0.046e-3, -19.15e-3, -0.00088, -0.00530], rtol=0.03)
    assert_allclose([R[0][0], R[0][1], R[0][5], R[1][1]], [-3.6, -3.3, -8.8, 6.85], atol=0.1)
    assert_allclose(-R[0], res.eR[0][0:6], atol=1e-9)               # element + support reaction = 0
    assert_allclose(-R[1], res.eR[0][6:12] + res.eR[1][0:6], atol=1e-9)


def test_element_math_closed_forms():
    """reference tests/python/test_np_stiffness.py:409-508."""
    from oracle.frame_oracle import bending_block, axial_block, rotation3
    L, E, I = 2.0, 3.0, 5.0
    k = E * I / L**3
    rigid = k * np.array([[12, 6 * L, -12, 6 * L], [6 * L, 4 * L**2, -6 * L, 2 * L**2],
                          [-12, -6 * L, 12, -6 * L], [6 * L, 2 * L**2, -6 * L, 4 * L**2]])
    assert_allclose(bending_block(L, E, I, 2), rigid, rtol=1e-14)
    assert_allclose(bending_block(L, E, I, 2, 0.0, 0.0), np.zeros((4, 4)), atol=0)
    pin_fix = k * np.array([[3, 0, -3, 3 * L], [0, 0, 0, 0], [-3, 0, 3, -3 * L], [3 * L, 0, -3 * L, 3 * L**2]])
    assert_allclose(bending_block(L, E, I, 2, 0.0, np.inf), pin_fix, rtol=1e-14, atol=1e-14)
    assert_allclose(axial_block(L, 2.0, E), E * 2.0 / L * np.array([[1, -1], [-1, 1]]))
    assert_allclose(axial_block(L, 2.0, E, 0.0, np.inf), np.zeros((2, 2)))
    assert_allclose(axial_block(L, 2.0, E, 4.0, 8.0), 1 / (L / (E * 2.0) + 1 / 4 + 1 / 8) * np.array([[1, -1], [-1, 1]]))
    assert_array_equal(rotation3([0, 0, 0], [3, 0, 0]), np.eye(3))
    assert_array_equal(rotation3([0, 0, 0], [0, 0, 2]), [[0, 0, -1], [0, 1, 0], [1, 0, 0]])     # Q3
