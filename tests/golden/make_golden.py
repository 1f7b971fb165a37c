"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the authoring container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports the unmodified reference numpy engine (``pyconmech`` from /root/reference/src,
with a no-op ``termcolor`` stub because that package is absent here), solves every case
listed below and freezes inputs + outputs as ``.npz``.  The model / load JSON inputs are
copied next to them (data, not source).  Nothing at test or bench time reads /root/reference.
"""
import io
import json
import os
import shutil
import sys
import types
import contextlib
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)

_tc = types.ModuleType('termcolor')
_tc.cprint = lambda *a, **k: None
_tc.colored = lambda s, *a, **k: s
sys.modules['termcolor'] = _tc
sys.path.insert(0, os.path.join(REF, 'src'))
warnings.simplefilter('ignore')

from pyconmech.frame_analysis.stiffness_checker import StiffnessChecker as RefChecker   # noqa: E402
from pyconmech.frame_analysis import io_base as rio                                     # noqa: E402

from conmech_b200 import synthetic as syn                                               # noqa: E402

MODELS = os.path.join(HERE, 'models')


def ref_model_from_ours(model):
    """Rebuild one of our synthetic Models as a reference Model through the JSON data layout."""
    return rio.Model.from_data(json.loads(json.dumps(model.to_data())))


def ref_loadcase_from_ours(lc):
    pls = [rio.PointLoad(pl.force, pl.moment, pl.node_ind) for pl in lc.point_loads]
    uls = [rio.UniformlyDistLoad(ul.q, ul.load, ul.elem_tags) for ul in lc.uniform_element_loads]
    g = rio.GravityLoad(lc.gravity_load.force) if lc.gravity_load is not None else None
    return rio.LoadCase(pls, uls, g)


def solve_record(sc, ids, keep_vectors=True):
    """One reference solve -> dict of order-free arrays (SURVEY §8c 'how the oracle must be wrapped')."""
    eng = sc._sc_ins
    nV, nE = eng.get_frame_stat()
    rec = {'ids': np.array(ids, dtype=np.int64)}
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            perm_out = eng.compute_dof_permutation(ids)
        except ValueError:
            perm_out = 'nosupport'
        try:
            flag = sc.solve(ids)
        except TypeError:          # Q10: bare False from compute_dof_permutation
            flag = False
            eng._has_stored_deformation = False
    rec['flag'] = np.array(bool(flag))
    rec['has_result'] = np.array(bool(sc.has_stored_result()))
    if isinstance(perm_out, tuple):
        Perm, (node_set, n_free, n_fixed, _) = perm_out
        rec['n_free'] = np.array(n_free)
        rec['n_fixed'] = np.array(n_fixed)
        if keep_vectors:
            rec['dof_map'] = Perm.tocsr().indices.astype(np.int32)
            ex = np.zeros(nV, dtype=bool)
            ex[list(node_set)] = True
            rec['node_exists'] = ex
    else:
        rec['n_free'] = np.array(-1)
        rec['n_fixed'] = np.array(-1)
    if sc.has_stored_result():
        _, nD, fR, eR = eng.get_solved_results()
        mt, mr, _, _ = eng.get_max_nodal_deformation()
        rec['max_trans'] = np.array(mt)
        rec['max_rot'] = np.array(mr)
        rec['compliance'] = np.array(eng.get_compliance())
        if keep_vectors:
            U = np.zeros(6 * nV)
            R = np.zeros(6 * nV)
            E = np.full((nE, 12), np.nan)
            for row in nD:
                U[6 * int(row[0]):6 * int(row[0]) + 6] = row[1:]
            for row in fR:
                R[6 * int(row[0]):6 * int(row[0]) + 6] = row[1:]
            for row in eR:
                E[int(row[0])] = row[1:]
            rec['U'], rec['R'], rec['eR'] = U, R, E
    return rec


def save_case(name, meta, records, extra=None):
    out = {'meta': np.array(json.dumps(meta))}
    out['n_solves'] = np.array(len(records))
    for i, rec in enumerate(records):
        for k, v in rec.items():
            out['s{}_{}'.format(i, k)] = v
    if extra:
        out.update(extra)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '{} solves, {:.0f} KB'.format(len(records), os.path.getsize(path) / 1024))


def element_tables(sc):
    eng = sc._sc_ins
    K = np.array(eng.get_element_stiffness_matrices())
    R = np.array(eng.get_element_local2global_rot_matrices())
    return {'K_G': K, 'R3': R[:, 0:3, 0:3].copy(), 'id_map': np.array(eng.get_element2dof_id_map(), dtype=np.int32)}


def copy_model(src, dst_name=None):
    os.makedirs(MODELS, exist_ok=True)
    dst = os.path.join(MODELS, dst_name or os.path.basename(src))
    shutil.copyfile(src, dst)
    os.chmod(dst, 0o644)
    return dst


def bundled_cases():
    td = os.path.join(REF, 'tests', 'test_data')
    for f in sorted(os.listdir(td)):
        copy_model(os.path.join(td, f))
    ex = os.path.join(REF, 'examples', 'scripts', 'cantilever_beam_truss')
    copy_model(os.path.join(ex, 'cantilever_beam_truss.json'))
    copy_model(os.path.join(ex, 'cantilever_beam_truss_loadcase.json'))

    tower_ok_sw = [1, 12, 5, 10, 4, 11, 13, 3, 2, 0, 7]
    tower_ok_pl = [0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15]
    tower_fail = [0, 4, 7, 8, 9]
    tower_mech = [0, 5, 6, 12, 15, 16]
    topopt_ok = [0, 1, 2, 3, 4, 5, 6, 8, 11, 14, 15, 19, 20, 21, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33,
                 34, 35, 36, 37, 38, 41, 42, 43, 44, 46, 51, 52, 53, 54, 55, 76, 88, 89, 90, 91, 92,
                 93, 94, 96, 97, 98, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113,
                 114, 115, 117, 119, 120, 121, 122, 123, 124, 125, 126, 127, 128, 129, 130, 131]
    topopt_fail = [1, 56, 59, 60, 66, 67, 68, 69, 83, 84, 102, 130]

    for stem, subsets in (('tower_3D', [[], tower_ok_sw, tower_ok_pl, tower_fail, tower_mech, [4, 5, 6, 7]]),
                          ('topopt-100', [[], topopt_ok, topopt_fail]),
                          ('tower_3D_broken_lines', [[]]),
                          ('four-frame', [[]]),
                          ('simple_frame', [[]])):
        path = os.path.join(td, stem + '.json')
        loads_path = os.path.join(td, stem + '_loads.json')
        for lc_name in ('self_weight', 'point_load', 'self_weight+point_load'):
            pls = []
            if 'point_load' in lc_name:
                with open(loads_path) as f:
                    raw = json.load(f)['loadcases']['0']
                pls = [rio.PointLoad.from_data(p) for p in raw['ploads']]
                if not pls:
                    continue
            grav = rio.GravityLoad([0, 0, -1]) if 'self_weight' in lc_name else None
            sc = RefChecker.from_json(path, checker_engine='numpy')
            sc.set_loads(rio.LoadCase(point_loads=pls, gravity_load=grav))
            sc.set_nodal_displacement_tol(trans_tol=1e-3)
            recs = [solve_record(sc, ids) for ids in subsets]
            extra = element_tables(sc) if lc_name == 'self_weight' else None
            if extra is not None:
                extra['q_gravity'] = np.array([sc._sc_ins._element_gravity_nload_list[e]
                                               for e in range(len(sc.elements))])
            save_case('{}__{}'.format(stem, lc_name),
                      {'model': stem + '.json', 'loads': stem + '_loads.json' if pls else None,
                       'gravity': [0, 0, -1] if grav else None, 'trans_tol': 1e-3}, recs, extra)

    # config 2: BFS construction sequences, gravity, tol 1e-3
    from conmech_b200.io_base import Model
    for stem, stride in (('tower_3D', 1), ('topopt-100', 6)):
        path = os.path.join(td, stem + '.json')
        order = syn.bfs_sequence(Model.from_json(path))
        sc = RefChecker.from_json(path, checker_engine='numpy')
        sc.set_loads(rio.LoadCase(gravity_load=rio.GravityLoad([0, 0, -1])))
        sc.set_nodal_displacement_tol(trans_tol=1e-3)
        recs = [solve_record(sc, order[:k], keep_vectors=(k % stride == 0 or k == len(order)))
                for k in range(1, len(order) + 1)]
        save_case('{}__bfs_prefixes'.format(stem),
                  {'model': stem + '.json', 'gravity': [0, 0, -1], 'trans_tol': 1e-3, 'order': order}, recs)

    # textbook cases driven through the facade
    sc = RefChecker.from_json(os.path.join(td, 'mgz_flat_beam.json'), checker_engine='numpy')
    sc.set_loads(rio.LoadCase.from_json(os.path.join(td, 'mgz_flat_beam_loads.json')))
    save_case('mgz_flat_beam__loads', {'model': 'mgz_flat_beam.json', 'loads': 'mgz_flat_beam_loads.json',
                                      'gravity': None, 'trans_tol': 1e-3}, [solve_record(sc, [])], element_tables(sc))
    sc = RefChecker.from_json(os.path.join(td, 'uniform_load_analytical_model.json'), checker_engine='numpy')
    sc.set_loads(rio.LoadCase.from_json(os.path.join(td, 'uniform_load_analytical_model_loads.json')))
    save_case('uniform_load_analytical_model__loads',
              {'model': 'uniform_load_analytical_model.json', 'loads': 'uniform_load_analytical_model_loads.json',
               'gravity': None, 'trans_tol': 1e-3}, [solve_record(sc, [])], element_tables(sc))

    # largest bundled model + the reference's own stored result
    sc = RefChecker.from_json(os.path.join(ex, 'cantilever_beam_truss.json'), checker_engine='numpy')
    with open(os.path.join(ex, 'cantilever_beam_truss_loadcase.json')) as f:
        sc.set_loads(rio.LoadCase.from_data(json.load(f)))       # flat LoadCase.to_data() layout
    sc.set_nodal_displacement_tol(trans_tol=1e-3)
    order = syn.bfs_sequence(Model.from_json(os.path.join(ex, 'cantilever_beam_truss.json')))
    recs = [solve_record(sc, [])] + [solve_record(sc, order[:k]) for k in (300, 1100, 1900)]
    with open(os.path.join(ex, 'cantilever_beam_truss_result.json')) as f:
        stored = json.load(f)
    nV = len(stored['nodal_displacement'])
    U = np.zeros(6 * nV)
    for k, v in stored['nodal_displacement'].items():
        U[6 * int(k):6 * int(k) + 6] = v
    R = np.zeros(6 * nV)
    for k, v in stored['support_reaction'].items():
        R[6 * int(k):6 * int(k) + 6] = v
    E = np.zeros((len(stored['element_reaction']), 12))
    for k, v in stored['element_reaction'].items():
        E[int(k)] = list(v[0]) + list(v[1])
    save_case('cantilever_beam_truss__loads',
              {'model': 'cantilever_beam_truss.json', 'loads': 'cantilever_beam_truss_loadcase.json',
               'gravity': None, 'trans_tol': 1e-3, 'order_prefixes': [0, 300, 1100, 1900]}, recs,
              {'stored_U': U, 'stored_R': R, 'stored_eR_presignflip': E,
               'stored_compliance': np.array(stored['elastic_energy']),
               'stored_max_trans': np.array(stored['max_trans'])})


def synthetic_cases():
    # config 3 shape, small and full size
    for dims, seeds, diag in (((4, 4, 4), range(6), False), ((4, 3, 4), range(6), True), ((7, 7, 8), range(3), False)):
        m = syn.grid_frame(*dims, diagonals=diag)
        sc = RefChecker(ref_model_from_ours(m), checker_engine='numpy')
        sc.set_loads(rio.LoadCase(gravity_load=rio.GravityLoad([0, 0, -1])))
        sc.set_nodal_displacement_tol(trans_tol=1e-3)
        subsets = [[]] + [syn.connected_subset(m, s) for s in seeds]
        recs = [solve_record(sc, ids) for ids in subsets]
        extra = element_tables(sc) if dims != (7, 7, 8) else None
        save_case('grid_{}x{}x{}{}__gravity'.format(*dims, '_diag' if diag else ''),
                  {'grid': list(dims), 'diagonals': diag, 'joints': False, 'gravity': [0, 0, -1],
                   'trans_tol': 1e-3, 'seeds': list(seeds)}, recs, extra)

    # config 4 shape: joints on horizontals, height-field subsets, four load cases
    dims = (5, 5, 5)
    m = syn.grid_frame(*dims, joints=True)
    lcs = syn.config4_loadcases(m)
    subsets = [[]] + [syn.heightfield_subset(m, *dims, seed=s) for s in range(8)]
    for li, lc in enumerate(lcs):
        sc = RefChecker(ref_model_from_ours(m), checker_engine='numpy')       # fresh checker per load case (Q6)
        sc.set_loads(ref_loadcase_from_ours(lc))
        sc.set_nodal_displacement_tol(trans_tol=1e-3)
        recs = [solve_record(sc, ids) for ids in subsets]
        extra = element_tables(sc) if li == 0 else None
        save_case('grid_5x5x5_joints__lc{}'.format(li),
                  {'grid': list(dims), 'diagonals': False, 'joints': True, 'loadcase': li,
                   'trans_tol': 1e-3, 'seeds': list(range(8))}, recs, extra)

    # free-form subsets with hinges: mostly singular (Q16) - flags / status only
    recs = []
    sc = RefChecker(ref_model_from_ours(m), checker_engine='numpy')
    sc.set_loads(rio.LoadCase(gravity_load=rio.GravityLoad([0, 0, -1])))
    for s in range(8):
        recs.append(solve_record(sc, syn.connected_subset(m, s), keep_vectors=False))
    save_case('grid_5x5x5_joints__freeform_flags',
              {'grid': list(dims), 'diagonals': False, 'joints': True, 'gravity': [0, 0, -1],
               'trans_tol': 1e-3, 'seeds': list(range(8))}, recs)


if __name__ == '__main__':
    bundled_cases()
    synthetic_cases()
